"""Sweep of the host entry point's pipeline knobs (NRQ_HOST_SLOTS x NRQ_HOST_CHUNK_BYTES x reads per call) and
the plain pinned-copy ceilings of the box.  python profiles/e2e_sweep.py > profiles/rNN_e2e_sweep.txt"""
import os, sys, time, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from nanoraw_b200 import _lib, synth
from nanoraw_b200.batch import ReadBatch
from nanoraw_b200.engine import ResquiggleParams

lib = _lib.load()
reads = synth.make_reads_parallel(2000, synth.config_standard())
base = ReadBatch.from_reads(reads)
pc = ResquiggleParams().to_c()

def tile(batch, k):
    if k == 1: return batch
    rs = reads * k
    return ReadBatch.from_reads(rs)

def run(hostb, label, steps=4):
    R, A = hostb.n_reads, hostb.total_align
    keep = []
    def pinned(arr):
        t = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory(); keep.append(t); return t
    hb = _lib.NrqBatch(); hb.n_reads = R
    for f in ReadBatch.FIELDS:
        setattr(hb, f, pinned(getattr(hostb, f)).data_ptr())
    outs = {'status': (R, np.int32), 'n_bases': (R, np.int32), 'n_indels': (R, np.int32), 'scale_values': (4*R, np.float64),
            'norm_mean': (A, np.float64), 'norm_stdev': (A, np.float64), 'ev_start': (A, np.int32), 'ev_length': (A, np.int32), 'ev_base': (A, np.uint8)}
    hr = _lib.NrqResults(); ho = {}
    for k, (n, dt) in outs.items():
        ho[k] = pinned(np.zeros(n, dtype=dt)); setattr(hr, k, ho[k].data_ptr())
    h = lib.nrq_create(0)
    for _ in range(2):
        _lib.check(lib.nrq_resquiggle_batch_host(h, C.byref(hb), C.byref(pc), C.byref(hr)), 'w')
    t0 = time.perf_counter()
    for _ in range(steps):
        _lib.check(lib.nrq_resquiggle_batch_host(h, C.byref(hb), C.byref(pc), C.byref(hr)), 's')
    ms = (time.perf_counter() - t0) * 1e3 / steps
    lib.nrq_destroy(h)
    ns = int(hostb.n_samples.sum())
    print('%-40s R=%6d %8.2f ms/call  %.3e samples/s' % (label, R, ms, ns / ms * 1e3), flush=True)

for R in (512, 2000, 8000):
    hostb = ReadBatch.from_reads((reads * 4)[:R])
    for slots in (2, 3, 4, 6, 8):
        for mb in (0, 16, 32, 64, 128):
            os.environ['NRQ_HOST_SLOTS'] = str(slots)
            if mb: os.environ['NRQ_HOST_CHUNK_BYTES'] = str(mb << 20)
            else: os.environ.pop('NRQ_HOST_CHUNK_BYTES', None)
            run(hostb, 'slots=%d chunk=%s' % (slots, ('%dMB' % mb) if mb else 'default'))
# pure copy ceilings
n = 512 << 20
a = torch.empty(n, dtype=torch.uint8).pin_memory(); b = torch.empty(n, dtype=torch.uint8).pin_memory()
da = torch.empty(n, dtype=torch.uint8, device='cuda'); db = torch.empty(n, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, rep=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(rep): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / rep
def h2d():
    with torch.cuda.stream(s1): da.copy_(a, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): b.copy_(db, non_blocking=True)
def both(): h2d(); d2h()
print('H2D GB/s', n / t(h2d) / 1e9); print('D2H GB/s', n / t(d2h) / 1e9); print('both GB/s total', 2 * n / t(both) / 1e9)
