#!/usr/bin/env python
"""K0 (indel scan) and K1 (median / MAD) read different inputs and are both far from saturating the SMs:
how long do they take one after the other vs side by side on two streams?"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from nanoraw_b200 import _lib, synth  # noqa: E402


def main():
    import torch
    from nanoraw_b200.batch import ReadBatch
    from nanoraw_b200.engine import ResquiggleEngine, ResquiggleParams, DeviceResults
    torch.cuda.set_device(0)
    engine = ResquiggleEngine()
    lib = engine.lib
    pc = ResquiggleParams().to_c()
    reads = synth.make_reads_parallel(1000, synth.config_standard(2000))
    db, _ = bench.build_device_batch(engine, ReadBatch.from_reads(reads), 100, 0)
    out = DeviceResults(db.n_reads, db.total_align, engine.device)
    dev = engine.device
    indel_off = torch.zeros(db.n_reads + 1, dtype=torch.int64, device=dev)
    cap = db.total_align // 4 + db.n_reads + 64
    indels = torch.empty(4 * cap, dtype=torch.int32, device=dev)
    status2 = torch.zeros(db.n_reads, dtype=torch.int32, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def k0(stream):
        _lib.check(lib.nrq_find_indels(C.byref(db.c), p(out.status), p(out.n_indels), p(out.n_bases), p(indel_off),
                                       p(indels), cap, p(out.ev_base), C.c_void_p(stream.cuda_stream)), 'k0')

    def k1(stream):
        _lib.check(lib.nrq_norm_stats(C.byref(db.c), C.byref(pc), p(status2), p(out.scale_values),
                                      C.c_void_p(stream.cuda_stream)), 'k1')

    def timed(fn, n=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    cur = torch.cuda.current_stream()

    def seq():
        k0(cur)
        k1(cur)

    def par():
        ev = torch.cuda.Event()
        ev.record(cur)
        s1.wait_event(ev)
        s2.wait_event(ev)
        k0(s1)
        k1(s2)
        for s in (s1, s2):
            d = torch.cuda.Event()
            d.record(s)
            cur.wait_event(d)

    print('K0 alone %.3f ms, K1 alone %.3f ms' % (timed(lambda: k0(cur)), timed(lambda: k1(cur))))
    print('one after the other %.3f ms, side by side %.3f ms' % (timed(seq), timed(par)))


if __name__ == '__main__':
    main()
