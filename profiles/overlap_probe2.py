#!/usr/bin/env python
"""Can the walk (k_resegment: a chain of trips to memory, a third of the issue slots idle) and the event table
(k_event_table: bound by instruction issue) share the SMs?  The walk is a persistent kernel that fills every SM, so
two streams alone do not overlap them; here the walk of read chunk c + 1 is launched with FEWER resident CTAs per SM
(NRQ_SEG_CTAS_PER_SM) and the event table of chunk c runs beside it on a second stream.

    python profiles/overlap_probe2.py --reads 100000 --chunks 4 8 --ctas 3 4 5

Stages through the C entry points: nrq_find_indels, nrq_norm_stats, nrq_resegment (debug_flags bit 4: speculative
searches only, bit 5: walk only), nrq_event_table.  Results are compared with the single-call path.
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from nanoraw_b200 import _lib, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reads', type=int, default=100000)
    ap.add_argument('--unique', type=int, default=1000)
    ap.add_argument('--chunks', type=int, nargs='+', default=[4, 8])
    ap.add_argument('--ctas', type=int, nargs='+', default=[3, 4, 5, 6])
    ap.add_argument('--steps', type=int, default=4)
    args = ap.parse_args()
    import torch
    from nanoraw_b200.batch import ReadBatch
    from nanoraw_b200.engine import ResquiggleEngine, ResquiggleParams, DeviceResults, _ptr
    torch.cuda.set_device(0)
    engine = ResquiggleEngine()
    lib = engine.lib
    dev = engine.device
    pc = ResquiggleParams().to_c()
    reads = synth.make_reads_parallel(args.unique, synth.config_standard(2000))
    hb = ReadBatch.from_reads(reads)
    db, n_samples = bench.build_device_batch(engine, hb, args.reads // args.unique, 0)
    R = db.n_reads
    out = DeviceResults(R, db.total_align, dev)
    ws = engine.workspace(R, db.total_align, pc)
    total = int(n_samples.sum().item())

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps

    os.environ.pop('NRQ_SEG_CTAS_PER_SM', None)
    ms = timed(lambda: engine.run_device(db, pc, out=out, workspace=ws))
    print('single call: %.2f ms/step  %.4g samples/s' % (ms, total / ms * 1e3), flush=True)
    ref = {k: getattr(out, k).clone() for k in ('status', 'new_segs', 'norm_mean', 'norm_stdev', 'ev_length')}
    n_ind = int(out.n_indels[:R].sum().item())

    # ---- the stages one by one, own workspace arrays
    cap = n_ind + 1024
    indel_off = torch.zeros(R + 1, dtype=torch.int64, device=dev)
    indels = torch.zeros(4 * cap, dtype=torch.int32, device=dev)
    groups = torch.zeros(4 * cap, dtype=torch.int32, device=dev)
    spec = torch.zeros(4 * cap, dtype=torch.int32, device=dev)
    big = torch.zeros(148 * 6 * 8 * (pc.big_region_cap + 1), dtype=torch.float64, device=dev)
    out2 = DeviceResults(R, db.total_align, dev)
    align_off = db.t['align_off']

    def sub(r0, r1):
        b = _lib.NrqBatch()
        b.n_reads = r1 - r0
        for name in ('raw', 'starts', 'read_align', 'genome_align'):
            setattr(b, name, db.t[name].data_ptr())
        for name in ('raw_off', 'starts_off', 'align_off'):
            setattr(b, name, db.t[name].data_ptr() + 8 * r0)
        b.read_start = db.t['read_start'].data_ptr() + 4 * r0
        return b

    def stage_calls(n_chunks, ctas):
        edges = np.linspace(0, R, n_chunks + 1).astype(int)
        sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
        counters = [torch.zeros(64, dtype=torch.int32, device=dev) for _ in range(n_chunks + 1)]
        p16 = ResquiggleParams().to_c(); p16.debug_flags = 16
        p32 = ResquiggleParams().to_c(); p32.debug_flags = 32
        whole = sub(0, R)

        def run():
            cur = torch.cuda.current_stream()
            st = C.c_void_p(cur.cuda_stream)
            _lib.check(lib.nrq_find_indels(C.byref(whole), _ptr(out2.status), _ptr(out2.n_indels), _ptr(out2.n_bases),
                                           _ptr(indel_off), _ptr(indels), cap, _ptr(out2.ev_base), st), 'find_indels')
            _lib.check(lib.nrq_norm_stats(C.byref(whole), C.byref(pc), _ptr(out2.status), _ptr(out2.scale_values), st), 'norm')
            # every speculative search first (one launch), then the chunks
            os.environ.pop('NRQ_SEG_CTAS_PER_SM', None)
            _lib.check(lib.nrq_resegment(C.byref(whole), C.byref(p16), _ptr(out2.scale_values), _ptr(indel_off), _ptr(indels),
                                         _ptr(out2.n_bases), _ptr(groups), _ptr(spec), cap, _ptr(big), big.numel(),
                                         _ptr(counters[n_chunks]), _ptr(out2.status), _ptr(out2.new_segs),
                                         _ptr(out2.region_samples), C.c_void_p(0), st), 'speculate')
            start = torch.cuda.Event()
            start.record(cur)
            sA.wait_event(start)
            sB.wait_event(start)
            for k in range(n_chunks):
                r0, r1 = int(edges[k]), int(edges[k + 1])
                b = sub(r0, r1)
                if k > 0 and ctas > 0:
                    os.environ['NRQ_SEG_CTAS_PER_SM'] = str(ctas)
                else:
                    os.environ.pop('NRQ_SEG_CTAS_PER_SM', None)
                _lib.check(lib.nrq_resegment(
                    C.byref(b), C.byref(p32), C.c_void_p(out2.scale_values.data_ptr() + 32 * r0),
                    C.c_void_p(indel_off.data_ptr() + 8 * r0), _ptr(indels), C.c_void_p(out2.n_bases.data_ptr() + 4 * r0),
                    _ptr(groups), _ptr(spec), cap, _ptr(big), big.numel(), _ptr(counters[k]),
                    C.c_void_p(out2.status.data_ptr() + 4 * r0), C.c_void_p(out2.new_segs.data_ptr() + 4 * r0),
                    C.c_void_p(out2.region_samples.data_ptr() + 4 * r0), C.c_void_p(0), C.c_void_p(sA.cuda_stream)), 'walk')
                ev = torch.cuda.Event()
                ev.record(sA)
                sB.wait_event(ev)
                _lib.check(lib.nrq_event_table(
                    C.byref(b), C.byref(pc), C.c_void_p(out2.scale_values.data_ptr() + 32 * r0),
                    C.c_void_p(out2.n_bases.data_ptr() + 4 * r0), C.c_void_p(out2.new_segs.data_ptr() + 4 * r0),
                    C.c_void_p(out2.status.data_ptr() + 4 * r0), _ptr(out2.norm_mean), _ptr(out2.norm_stdev),
                    _ptr(out2.ev_start), _ptr(out2.ev_length), _ptr(out2.ev_base), C.c_void_p(sB.cuda_stream)), 'events')
            os.environ.pop('NRQ_SEG_CTAS_PER_SM', None)
            for s_ in (sA, sB):
                done = torch.cuda.Event()
                done.record(s_)
                cur.wait_event(done)
        return run

    ao = align_off[:-1]
    edge = torch.zeros(db.total_align + 1, dtype=torch.int32, device=dev)
    edge.index_add_(0, ao, torch.ones_like(ao, dtype=torch.int32))
    edge.index_add_(0, ao + out.n_bases[:R].long(), -torch.ones_like(ao, dtype=torch.int32))
    valid = torch.cumsum(edge, 0)[:db.total_align] > 0
    for n_chunks in args.chunks:
        for ctas in args.ctas:
            run = stage_calls(n_chunks, ctas)
            out2.norm_mean.fill_(-7.0)
            ms = timed(run)
            same = bool(torch.equal(out2.status[:R], ref['status'][:R]) and
                        torch.equal(out2.norm_mean[:db.total_align][valid], ref['norm_mean'][:db.total_align][valid]) and
                        torch.equal(out2.norm_stdev[:db.total_align][valid], ref['norm_stdev'][:db.total_align][valid]))
            print('chunks=%2d walk CTAs/SM beside the event table=%d: %.2f ms/step  %.4g samples/s  identical=%s' % (
                n_chunks, ctas, ms, total / ms * 1e3, same), flush=True)


if __name__ == '__main__':
    main()
