#!/usr/bin/env python
"""Does running the four stages of different read chunks side by side (several CUDA streams) beat one
whole-batch launch per stage?  K2 is instruction-issue bound and barely touches DRAM, K0/K1 are the
opposite, and every kernel has a tail: chunks on separate streams let them fill each other's gaps.

    python profiles/overlap_probe.py --reads 100000 --chunks 8 --streams 4

Builds the same device batch as bench.py, then times (CUDA events, whole batch) the single-call path and
the chunked multi-stream path through the same C entry point (nrq_resquiggle_batch on sub-ranges of reads).
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


def sub_batch(db, out, r0, r1):
    """nrq_batch / nrq_results structs for reads [r0, r1) of a device batch (offset arrays hold absolute
    offsets, so a sub-range is a pointer shift of the per-read arrays)"""
    b = _lib.NrqBatch()
    b.n_reads = r1 - r0
    for name in ('raw', 'starts', 'read_align', 'genome_align'):
        setattr(b, name, db.t[name].data_ptr())
    for name in ('raw_off', 'starts_off', 'align_off'):
        setattr(b, name, db.t[name].data_ptr() + 8 * r0)
    b.read_start = db.t['read_start'].data_ptr() + 4 * r0
    o = _lib.NrqResults()
    for name in ('norm_mean', 'norm_stdev', 'ev_start', 'ev_length', 'ev_base'):
        setattr(o, name, getattr(out, name).data_ptr())
    for name in ('status', 'n_bases', 'n_indels', 'region_samples'):
        setattr(o, name, getattr(out, name).data_ptr() + 4 * r0)
    o.scale_values = out.scale_values.data_ptr() + 32 * r0
    o.new_segs = out.new_segs.data_ptr() + 4 * r0  # a read's slot is align_off[r] + r
    return b, o


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reads', type=int, default=100000)
    ap.add_argument('--unique', type=int, default=1000)
    ap.add_argument('--chunks', type=int, nargs='+', default=[4, 8, 16])
    ap.add_argument('--streams', type=int, nargs='+', default=[2, 3, 4])
    ap.add_argument('--steps', type=int, default=4)
    args = ap.parse_args()
    import torch
    from nanoraw_b200.batch import ReadBatch
    from nanoraw_b200.engine import ResquiggleEngine, ResquiggleParams, DeviceResults
    torch.cuda.set_device(0)
    engine = ResquiggleEngine()
    lib = engine.lib
    pc = ResquiggleParams().to_c()
    reads = synth.make_reads_parallel(args.unique, synth.config_standard(2000))
    hb = ReadBatch.from_reads(reads)
    reps = args.reads // args.unique
    db, n_samples = bench.build_device_batch(engine, hb, reps, 0)
    out = DeviceResults(db.n_reads, db.total_align, engine.device)
    ws = engine.workspace(db.n_reads, db.total_align, pc)
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

    ms = timed(lambda: engine.run_device(db, pc, out=out, workspace=ws))
    ref_status = out.status.clone()
    ref_mean = out.norm_mean.clone()
    ref_segs = out.new_segs.clone()
    # event rows that hold results: [align_off[r], align_off[r] + n_bases[r])
    ao = db.t['align_off'][:-1]
    edge = torch.zeros(db.total_align + 1, dtype=torch.int32, device=engine.device)
    edge.index_add_(0, ao, torch.ones_like(ao, dtype=torch.int32))
    edge.index_add_(0, ao + out.n_bases[:db.n_reads].long(), -torch.ones_like(ao, dtype=torch.int32))
    valid = torch.cumsum(edge, 0)[:db.total_align] > 0
    print('single call: %.2f ms/step  %.4g samples/s' % (ms, total / ms * 1e3), flush=True)

    align_off = db.t['align_off'].cpu().numpy()
    for n_chunks in args.chunks:
        edges = np.linspace(0, db.n_reads, n_chunks + 1).astype(int)
        for n_streams in args.streams:
            streams = [torch.cuda.Stream() for _ in range(n_streams)]
            wss = []
            for k in range(n_chunks):
                r0, r1 = int(edges[k]), int(edges[k + 1])
                need = int(lib.nrq_workspace_bytes(r1 - r0, int(align_off[r1] - align_off[r0]), C.byref(pc)))
                wss.append(torch.empty(need, dtype=torch.uint8, device=engine.device))
            subs = [sub_batch(db, out, int(edges[k]), int(edges[k + 1])) for k in range(n_chunks)]

            def run_chunked():
                cur = torch.cuda.current_stream()
                start = torch.cuda.Event()
                start.record(cur)
                for k in range(n_chunks):
                    st = streams[k % n_streams]
                    if k < n_streams:
                        st.wait_event(start)
                    b, o = subs[k]
                    rc = lib.nrq_resquiggle_batch(C.byref(b), C.byref(pc), C.byref(o), C.c_void_p(wss[k].data_ptr()),
                                                  wss[k].numel(), C.c_void_p(st.cuda_stream))
                    _lib.check(rc, 'chunk')
                for st in streams:
                    done = torch.cuda.Event()
                    done.record(st)
                    cur.wait_event(done)

            out.norm_mean.fill_(-7.0)
            out.new_segs.fill_(-7)
            ms = timed(run_chunked)
            seg_ok = ((out.new_segs == ref_segs) | (out.new_segs == -7)).all() and (out.new_segs != -7).sum() > 0
            same = bool(torch.equal(out.status, ref_status) and seg_ok and
                        torch.equal(out.norm_mean[:db.total_align][valid], ref_mean[:db.total_align][valid]))
            print('chunks=%2d streams=%d: %.2f ms/step  %.4g samples/s  identical=%s' % (
                n_chunks, n_streams, ms, total / ms * 1e3, same), flush=True)
            del wss


if __name__ == '__main__':
    main()
