"""Device pipeline: ReadBatch -> libnrq kernels -> per-read results.

PyTorch owns device memory and the CUDA stream; every computation is a libnrq kernel
reached through the C ABI in include/nrq.h.  No CPU fallback exists: constructing an
engine without a CUDA device or without libnrq.so raises.
"""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib
from .batch import ReadBatch

EVENT_DTYPE = np.dtype([('norm_mean', '<f8'), ('norm_stdev', '<f8'),
                        ('start', '<u4'), ('length', '<u4'), ('base', 'S1')])


@dataclass
class ResquiggleParams:
    """Batch-wide options = the CLI-level arguments of resquiggle_read (resquiggle.py:318-323)."""
    norm_type: str = 'median'            # 'median' | 'none' | 'pA_raw' | 'pA' | 'given'
    outlier_thresh: Optional[float] = 5  # None disables clipping (resquiggle.py:1159-1160)
    compute_sd: bool = True
    num_cpts_limit: Optional[int] = None
    timeout: Optional[float] = None      # seconds, measured on the device clock
    min_event_obs: int = 4
    given_limits: bool = False
    big_region_cap: int = 16384
    exact_stats: bool = False            # event mean/sd in numpy's summation order (bit-identical, slower)
    long_read_parts: int = 0             # nrq_params.long_read_parts (0: the host entry point decides per chunk)
    debug_flags: int = 0

    def to_c(self):
        p = _lib.NrqParams()
        p.norm_type = _lib.NORM_MEDIAN if self.norm_type == 'median' else _lib.NORM_GIVEN
        p.use_outlier = 0 if self.outlier_thresh is None else 1
        p.outlier_thresh = 0.0 if self.outlier_thresh is None else float(self.outlier_thresh)
        p.given_limits = int(self.given_limits)
        p.compute_sd = int(self.compute_sd)
        p.num_cpts_limit = -1 if self.num_cpts_limit is None else int(self.num_cpts_limit)
        p.min_seg_len = int(self.min_event_obs)
        p.timeout_ns = 0 if self.timeout is None else int(self.timeout * 1e9)
        p.big_region_cap = int(self.big_region_cap)
        p.long_read_parts = int(self.long_read_parts)
        p.debug_flags = int(self.debug_flags) | (2 if self.exact_stats else 0)
        return p


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class DeviceBatch(object):
    """A ReadBatch resident in HBM (torch tensors) plus the nrq_batch struct pointing at it."""
    OPTIONAL = ('scale_values_in', 'norm_signal', 'norm_off')

    def __init__(self, batch: ReadBatch, device, non_blocking=False, pin=False):
        import torch
        self.n_reads = batch.n_reads
        self.total_align = batch.total_align
        self.device = device
        self.host = batch
        self.t = {}
        for name in ReadBatch.FIELDS + self.OPTIONAL:
            arr = getattr(batch, name, None)
            if arr is None:
                self.t[name] = None
                continue
            ht = torch.from_numpy(np.ascontiguousarray(arr))
            if pin:
                ht = ht.pin_memory()
            self.t[name] = ht.to(device, non_blocking=non_blocking)
        self.c = self._struct()

    @classmethod
    def from_tensors(cls, tensors, n_reads, total_align, device, host=None):
        """Wrap tensors that already live on `device` (bench.py builds big batches there)."""
        self = cls.__new__(cls)
        self.n_reads = n_reads
        self.total_align = total_align
        self.device = device
        self.host = host
        self.t = {name: tensors.get(name) for name in ReadBatch.FIELDS + self.OPTIONAL}
        self.c = self._struct()
        return self

    def _struct(self):
        b = _lib.NrqBatch()
        b.n_reads = self.n_reads
        for name in ReadBatch.FIELDS + self.OPTIONAL:
            setattr(b, name, _ptr(self.t[name]))
        return b

    def device_bytes(self):
        return sum(t.numel() * t.element_size() for t in self.t.values() if t is not None)


class DeviceResults(object):
    """Output arrays of nrq_resquiggle_batch in HBM."""

    def __init__(self, n_reads, total_align, device):
        import torch
        R, A = n_reads, total_align
        mk = lambda n, dt: torch.empty(max(int(n), 1), dtype=dt, device=device)
        self.status = mk(R, torch.int32)
        self.n_bases = mk(R, torch.int32)
        self.n_indels = mk(R, torch.int32)
        self.scale_values = mk(4 * R, torch.float64)
        self.new_segs = mk(A + R, torch.int32)
        self.norm_mean = mk(A, torch.float64)
        self.norm_stdev = mk(A, torch.float64)
        self.ev_start = mk(A, torch.int32)    # bit pattern of uint32
        self.ev_length = mk(A, torch.int32)
        self.ev_base = mk(A, torch.uint8)
        self.region_samples = mk(R, torch.int32)
        c = _lib.NrqResults()
        for name, _ in _lib.NrqResults._fields_:
            setattr(c, name, _ptr(getattr(self, name)))
        self.c = c

    NAMES = tuple(n for n, _ in _lib.NrqResults._fields_)

    def device_bytes(self):
        return sum(getattr(self, n).numel() * getattr(self, n).element_size() for n in self.NAMES)


class HostResults(object):
    """Results of a batch on the host, with per-read accessors shaped like the reference's
    outputs (new_segs: resquiggle.py:372; event table: resquiggle.py:66-70)."""

    def __init__(self, align_off, arrays):
        self.align_off = np.asarray(align_off)
        for k, v in arrays.items():
            setattr(self, k, v)
        self.scale_values = self.scale_values.reshape(-1, 4)

    @property
    def n_reads(self):
        return len(self.status)

    def ok(self, r):
        return int(self.status[r]) == _lib.ST_OK

    def error(self, r):
        return _lib.status_message(self.status[r])

    def new_segs_of(self, r):
        o = int(self.align_off[r]) + r
        return self.new_segs[o:o + int(self.n_bases[r]) + 1].astype(np.int64)

    def scale_values_of(self, r):
        """(shift, scale, lower_lim, upper_lim); None where the reference has None."""
        return tuple(None if np.isnan(v) else float(v) for v in self.scale_values[r])

    def events_of(self, r):
        o, n = int(self.align_off[r]), int(self.n_bases[r])
        ev = np.empty(n, dtype=EVENT_DTYPE)
        ev['norm_mean'] = self.norm_mean[o:o + n]
        ev['norm_stdev'] = self.norm_stdev[o:o + n]
        ev['start'] = self.ev_start[o:o + n]
        ev['length'] = self.ev_length[o:o + n]
        ev['base'] = self.ev_base[o:o + n].view('S1')
        return ev


class ResquiggleEngine(object):
    """Runs the a1..a7 hot path for batches of reads on one GPU."""

    def __init__(self, device=None):
        import torch
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.NrqError('no CUDA device: the re-squiggle path has no CPU fallback')
        self.torch = torch
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
        self._ws = None

    # ---- plumbing
    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def workspace(self, n_reads, total_align, params_c, min_bytes=0):
        need = max(int(self.lib.nrq_workspace_bytes(n_reads, total_align, C.byref(params_c))), min_bytes)
        if need < 0:
            raise _lib.NrqError(self.lib.nrq_last_error().decode())
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = self.torch.empty(need, dtype=self.torch.uint8, device=self.device)
        return self._ws

    def upload(self, batch, pin=False):
        with self.torch.cuda.device(self.device):
            return DeviceBatch(batch, self.device, pin=pin)

    def alloc_results(self, dbatch):
        return DeviceResults(dbatch.n_reads, dbatch.total_align, self.device)

    # ---- the whole path, device resident (kernel-only timing uses this)
    def run_device(self, dbatch, params, out=None, workspace=None):
        pc = params.to_c() if isinstance(params, ResquiggleParams) else params
        with self.torch.cuda.device(self.device):
            out = out or self.alloc_results(dbatch)
            ws = workspace if workspace is not None else self.workspace(
                dbatch.n_reads, dbatch.total_align, pc)
            rc = self.lib.nrq_resquiggle_batch(
                C.byref(dbatch.c), C.byref(pc), C.byref(out.c), _ptr(ws), ws.numel(), self._stream())
        _lib.check(rc, 'nrq_resquiggle_batch')
        return out

    def run_device_profiled(self, dbatch, params, out, workspace):
        """run_device, plus the device time (ms) of {indel scan, order statistics, speculative searches, walk,
        event table} (CUDA events on the launching stream); returns after the batch has finished."""
        pc = params.to_c() if isinstance(params, ResquiggleParams) else params
        ms = (C.c_float * 5)()
        with self.torch.cuda.device(self.device):
            rc = self.lib.nrq_resquiggle_batch_profile(
                C.byref(dbatch.c), C.byref(pc), C.byref(out.c), _ptr(workspace), workspace.numel(),
                self._stream(), ms)
        _lib.check(rc, 'nrq_resquiggle_batch_profile')
        return list(ms)

    def download(self, dbatch, out):
        arrays = {}
        for name in DeviceResults.NAMES:
            a = getattr(out, name).cpu().numpy()
            if name in ('ev_start', 'ev_length'):
                a = a.view(np.uint32)
            arrays[name] = a
        R = dbatch.n_reads
        for name in ('status', 'n_bases', 'n_indels', 'region_samples'):
            arrays[name] = arrays[name][:R]
        arrays['scale_values'] = arrays['scale_values'][:4 * R]
        return HostResults(dbatch.host.align_off, arrays)

    # ---- host buffers in, host buffers out: the C-ABI host entry point (chunked, multi-stream)
    def _host_handle(self):
        if getattr(self, '_handle', None) is None:
            import atexit
            h = self.lib.nrq_create(self.device.index or 0)
            if not h:
                raise _lib.NrqError('nrq_create failed: ' + self.lib.nrq_last_error().decode())
            self._handle = h
            atexit.register(self._release_handle)
        return self._handle

    def _release_handle(self):
        h, self._handle = getattr(self, '_handle', None), None
        if h:
            self.lib.nrq_destroy(h)

    def run_host(self, src, params, skip=()):
        """Run a batch whose CSR arrays live in host memory (a ReadBatch or a PinnedBatcher) through
        nrq_resquiggle_batch_host.  Results are numpy arrays: views of the batcher's pinned output
        buffers (valid until its next reset) or fresh arrays for a ReadBatch.  `skip`: result arrays that are not
        to be copied back from the device (NULL in nrq_results), e.g. new_segs and ev_start for a caller that only
        hands the event table on (start is the running sum of length)."""
        R, A = src.n_reads, src.total_align
        pc = params.to_c() if isinstance(params, ResquiggleParams) else params
        hb = _lib.NrqBatch()
        hb.n_reads = R
        keep = []
        for name in ReadBatch.FIELDS:
            arr = np.ascontiguousarray(getattr(src, name))
            keep.append(arr)
            setattr(hb, name, arr.ctypes.data)
        sv_in = getattr(src, 'scale_values_in', None)
        if sv_in is not None and (getattr(src, 'has_given', True)):
            sv_in = np.ascontiguousarray(sv_in, dtype=np.float64)
            keep.append(sv_in)
            hb.scale_values_in = sv_in.ctypes.data
        outs = {}
        shapes = {'status': (R, np.int32), 'n_bases': (R, np.int32), 'n_indels': (R, np.int32),
                  'region_samples': (R, np.int32), 'scale_values': (4 * R, np.float64),
                  'new_segs': (A + R, np.int32), 'norm_mean': (A, np.float64),
                  'norm_stdev': (A, np.float64), 'ev_start': (A, np.uint32), 'ev_length': (A, np.uint32),
                  'ev_base': (A, np.uint8)}
        hr = _lib.NrqResults()
        own = hasattr(src, 'reset')  # a PinnedBatcher brings its own output buffers
        for name, (n, dt) in shapes.items():
            if name in skip:
                outs[name] = None
                continue
            arr = getattr(src, name)[:max(n, 1)] if own else np.empty(max(n, 1), dtype=dt)
            outs[name] = arr[:n]
            setattr(hr, name, arr.ctypes.data)
        if R:
            rc = self.lib.nrq_resquiggle_batch_host(self._host_handle(), C.byref(hb), C.byref(pc), C.byref(hr))
            _lib.check(rc, 'nrq_resquiggle_batch_host')
        res = HostResults(np.asarray(src.align_off[:R + 1]), outs)
        # the C structs (host pointers) of this batch, for callers that hand the results on through the C ABI
        # (libnrqio's chunk packing); `keep` pins the arrays they point at
        res.c_batch, res.c_results, res._keep = hb, hr, keep
        return res

    def run(self, batch, params):
        """Host arrays in, host arrays out (ReadBatch -> HostResults)."""
        return self.run_host(batch, params)

    def run_via_torch(self, batch, params):
        """Same through torch-managed device tensors and nrq_resquiggle_batch.  Reads the indel
        workspace could not hold are re-run once with a workspace sized for the worst case."""
        db = self.upload(batch)
        pc = params.to_c()
        out = self.run_device(db, pc)
        res = self.download(db, out)
        if (res.status == _lib.ST_WORKSPACE).any():
            ws = self.workspace(db.n_reads, db.total_align, pc,
                                min_bytes=32 * (db.total_align + db.n_reads) + (1 << 20) +
                                int(self.lib.nrq_workspace_bytes(db.n_reads, 0, C.byref(pc))))
            out = self.run_device(db, pc, out=out, workspace=ws)
            res = self.download(db, out)
        return res

    # ---- single stages (tests, normalize_raw_signal / get_indel_groups call surfaces)
    def find_indels(self, dbatch):
        """-> (status, n_indels, n_bases, indel_off, indels[I,4]) on the host."""
        t = self.torch
        R = dbatch.n_reads
        with t.cuda.device(self.device):
            status = t.empty(R, dtype=t.int32, device=self.device)
            n_ind = t.empty(R, dtype=t.int32, device=self.device)
            n_bases = t.empty(R, dtype=t.int32, device=self.device)
            off = t.empty(R + 1, dtype=t.int64, device=self.device)
            rc = self.lib.nrq_find_indels(C.byref(dbatch.c), _ptr(status), _ptr(n_ind), _ptr(n_bases),
                                          _ptr(off), C.c_void_p(0), 0, C.c_void_p(0), self._stream())
            _lib.check(rc, 'nrq_find_indels')
            total = int(off[-1].item())
            indels = t.empty(4 * max(total, 1), dtype=t.int32, device=self.device)
            bases = t.empty(max(dbatch.total_align, 1), dtype=t.uint8, device=self.device)
            rc = self.lib.nrq_find_indels(C.byref(dbatch.c), _ptr(status), _ptr(n_ind), _ptr(n_bases),
                                          _ptr(off), _ptr(indels), total, _ptr(bases), self._stream())
            _lib.check(rc, 'nrq_find_indels')
        return (status.cpu().numpy(), n_ind.cpu().numpy(), n_bases.cpu().numpy(), off.cpu().numpy(),
                indels.cpu().numpy()[:4 * total].reshape(-1, 4), bases.cpu().numpy())

    def norm_stats(self, dbatch, params, status=None):
        t = self.torch
        R = dbatch.n_reads
        pc = params.to_c()
        with t.cuda.device(self.device):
            if status is None:
                status = t.zeros(R, dtype=t.int32, device=self.device)
            sv = t.empty(4 * R, dtype=t.float64, device=self.device)
            rc = self.lib.nrq_norm_stats(C.byref(dbatch.c), C.byref(pc), _ptr(status), _ptr(sv), self._stream())
            _lib.check(rc, 'nrq_norm_stats')
        return status, sv

    def normalize_signal(self, dbatch, params):
        """-> (status, scale_values[R,4], [float64 normalised signal per read]) on the host."""
        t = self.torch
        status, sv = self.norm_stats(dbatch, params)
        n_samp = dbatch.host.n_samples
        off = np.zeros(dbatch.n_reads + 1, dtype=np.int64)
        np.cumsum(n_samp, out=off[1:])
        with t.cuda.device(self.device):
            d_off = t.from_numpy(off).to(self.device)
            sig = t.full((max(int(off[-1]), 1),), float('nan'), dtype=t.float64, device=self.device)
            rc = self.lib.nrq_normalize_signal(C.byref(dbatch.c), _ptr(sv), _ptr(status), _ptr(d_off),
                                               _ptr(sig), self._stream())
            _lib.check(rc, 'nrq_normalize_signal')
        st = status.cpu().numpy()
        svh = sv.cpu().numpy().reshape(-1, 4)
        h = sig.cpu().numpy()
        raw_len = np.diff(dbatch.host.raw_off) - dbatch.host.read_start
        sigs = [h[off[r]:off[r] + min(int(n_samp[r]), max(int(raw_len[r]), 0))]
                for r in range(dbatch.n_reads)]
        return st, svh, sigs

    def indel_groups(self, read_align, genome_align, align_segs, norm_signal, min_seg_len=4,
                     timeout=None, num_cpts_limit=None):
        """get_indel_groups (resquiggle.py:139-316) for ONE read whose normalised float64 signal is
        given.  -> (status, [(group_start, group_stop, cpts, indels)]) with indels as (start, end, diff)."""
        t = self.torch
        sig = np.ascontiguousarray(norm_signal, dtype=np.float64)
        align_segs = np.asarray(align_segs)
        if len(align_segs) and int(align_segs[-1]) > len(sig):
            # the reference slices raw_signal[a:b] and would silently run on a truncated region; the kernel indexes
            # the signal directly, so a signal shorter than the last boundary is refused up front with the message
            # resquiggle_read gives for it (resquiggle.py:379-381)
            return _lib.ST_PAST_END, []
        batch = ReadBatch.from_arrays([np.zeros(1, np.int16)], [0], [align_segs], [read_align],
                                      [genome_align])
        batch.norm_signal = sig
        batch.norm_off = np.array([0, len(sig)], dtype=np.int64)
        db = self.upload(batch)
        status, n_ind, n_bases, off, indels, _ = self.find_indels(db)
        params = ResquiggleParams(num_cpts_limit=num_cpts_limit, timeout=timeout,
                                  min_event_obs=min_seg_len, outlier_thresh=None)
        pc = params.to_c()
        total = max(int(off[-1]), 1)
        with t.cuda.device(self.device):
            d_status = t.from_numpy(status.copy()).to(self.device)
            d_off = t.from_numpy(off).to(self.device)
            d_ind = t.from_numpy(np.ascontiguousarray(indels.reshape(-1))).to(self.device) \
                if len(indels) else t.zeros(4, dtype=t.int32, device=self.device)
            d_nb = t.from_numpy(n_bases).to(self.device)
            groups = t.zeros(4 * total, dtype=t.int32, device=self.device)
            ws = self.workspace(1, batch.total_align, pc)
            big = t.empty(ws.numel() // 8, dtype=t.float64, device=self.device)
            counter = t.zeros(8, dtype=t.int32, device=self.device)
            sv = t.tensor([0.0, 1.0, float('nan'), float('nan')], dtype=t.float64, device=self.device)
            segs = t.zeros(batch.total_align + 2, dtype=t.int32, device=self.device)
            ng = t.zeros(1, dtype=t.int32, device=self.device)
            rc = self.lib.nrq_resegment(
                C.byref(db.c), C.byref(pc), _ptr(sv), _ptr(d_off), _ptr(d_ind), _ptr(d_nb), _ptr(groups),
                C.c_void_p(0), 0, _ptr(big), big.numel(), _ptr(counter), _ptr(d_status), _ptr(segs), C.c_void_p(0),
                _ptr(ng), self._stream())
            _lib.check(rc, 'nrq_resegment')
        st = int(d_status.cpu()[0])
        out = []
        if st == _lib.ST_OK:
            g = groups.cpu().numpy().reshape(-1, 4)[:int(ng.cpu()[0])]
            new_segs = segs.cpu().numpy()
            dpre = indels[:, 3] if len(indels) else np.zeros(0, np.int32)
            for gi, (g_start, g_stop, lo, _maxend) in enumerate(g):
                hi = int(g[gi + 1][2]) if gi + 1 < len(g) else len(indels)
                grp = indels[lo:hi]
                k = int(grp[:, 2].sum()) + g_stop - g_start - 1
                first = g_start + int(dpre[lo]) + 1
                out.append((int(g_start), int(g_stop), new_segs[first:first + k].astype(np.int64).tolist(),
                            [tuple(int(v) for v in row[:3]) for row in grp]))
        return st, out

    def genome_aggregate(self, starts, lengths, reverse, means, sds, ev_lengths, track_len):
        """One (chromosome, strand) track of get_coverage / get_base_means / get_base_sds / get_base_lengths
        (nanoraw_helper.py:298-444).  starts / lengths / reverse: per read; means / sds / ev_lengths: per-read
        event columns (lists of arrays).  -> (coverage int32[L], mean of means, mean of sds, mean of lengths)."""
        t = self.torch
        L = int(track_len)
        order = np.argsort(np.asarray(starts), kind='stable')
        n_reads = len(order)
        st = np.asarray(starts, dtype=np.int32)[order]
        ln = np.asarray(lengths, dtype=np.int32)[order]
        rv = np.asarray(reverse, dtype=np.uint8)[order]
        off = np.zeros(n_reads + 1, dtype=np.int64)
        np.cumsum(ln, out=off[1:])
        cat = lambda cols, dt: (np.concatenate([np.asarray(cols[i], dtype=dt) for i in order])
                                if n_reads else np.zeros(1, dt))
        with t.cuda.device(self.device):
            dev = lambda a: t.from_numpy(np.ascontiguousarray(a)).to(self.device)
            d_st, d_ln, d_rv, d_off = dev(st if n_reads else np.zeros(1, np.int32)), \
                dev(ln if n_reads else np.zeros(1, np.int32)), dev(rv if n_reads else np.zeros(1, np.uint8)), \
                dev(off[:-1] if n_reads else np.zeros(1, np.int64))
            d_m, d_s = dev(cat(means, np.float64)), dev(cat(sds, np.float64))
            d_l = dev(cat(ev_lengths, np.uint32).view(np.int32))
            cov = t.empty(max(L, 1), dtype=t.int32, device=self.device)
            o_m = t.empty(max(L, 1), dtype=t.float64, device=self.device)
            o_s = t.empty(max(L, 1), dtype=t.float64, device=self.device)
            o_l = t.empty(max(L, 1), dtype=t.float64, device=self.device)
            rc = self.lib.nrq_genome_aggregate(
                _ptr(d_st), _ptr(d_ln), _ptr(d_off), _ptr(d_rv), n_reads, int(ln.max()) if n_reads else 0, L,
                _ptr(d_m), _ptr(d_s), _ptr(d_l), _ptr(cov), _ptr(o_m), _ptr(o_s), _ptr(o_l), self._stream())
            _lib.check(rc, 'nrq_genome_aggregate')
        return (cov.cpu().numpy()[:L], o_m.cpu().numpy()[:L], o_s.cpu().numpy()[:L], o_l.cpu().numpy()[:L])

    def genome_event_lists(self, starts, lengths, reverse, means, track_len):
        """get_reads_events (nanoraw_helper.py:446-484) for one (chromosome, strand) track: the event means of every
        read covering a position.  -> (coverage int64[L], pos_off int64[L + 1], values float64[pos_off[-1]])."""
        t = self.torch
        L = int(track_len)
        order = np.argsort(np.asarray(starts), kind='stable')
        n_reads = len(order)
        if n_reads == 0 or L <= 0:
            return np.zeros(max(L, 0), np.int64), np.zeros(max(L, 0) + 1, np.int64), np.zeros(0)
        st = np.asarray(starts, dtype=np.int32)[order]
        ln = np.asarray(lengths, dtype=np.int32)[order]
        rv = np.asarray(reverse, dtype=np.uint8)[order]
        off = np.zeros(n_reads + 1, dtype=np.int64)
        np.cumsum(ln, out=off[1:])
        cat = np.concatenate([np.asarray(means[i], dtype=np.float64) for i in order])
        zeros = np.zeros(len(cat))
        with t.cuda.device(self.device):
            dev = lambda a: t.from_numpy(np.ascontiguousarray(a)).to(self.device)
            d_st, d_ln, d_rv, d_off, d_m = dev(st), dev(ln), dev(rv), dev(off[:-1]), dev(cat)
            d_z = dev(zeros)
            d_zl = t.zeros(len(cat), dtype=t.int32, device=self.device)
            cov = t.empty(L, dtype=t.int32, device=self.device)
            scratch = [t.empty(L, dtype=t.float64, device=self.device) for _ in range(3)]
            rc = self.lib.nrq_genome_aggregate(
                _ptr(d_st), _ptr(d_ln), _ptr(d_off), _ptr(d_rv), n_reads, int(ln.max()), L, _ptr(d_m), _ptr(d_z),
                _ptr(d_zl), _ptr(cov), _ptr(scratch[0]), _ptr(scratch[1]), _ptr(scratch[2]), self._stream())
            _lib.check(rc, 'nrq_genome_aggregate')
            pos_off = t.zeros(L + 1, dtype=t.int64, device=self.device)
            pos_off[1:] = t.cumsum(cov.to(t.int64), 0)  # offsets only: plumbing, the lists are filled by the kernel
            total = int(pos_off[-1].item())
            values = t.empty(max(total, 1), dtype=t.float64, device=self.device)
            rc = self.lib.nrq_genome_event_lists(
                _ptr(d_st), _ptr(d_ln), _ptr(d_off), _ptr(d_rv), n_reads, int(ln.max()), L, _ptr(d_m),
                _ptr(pos_off), _ptr(values), self._stream())
            _lib.check(rc, 'nrq_genome_event_lists')
        return cov.cpu().numpy().astype(np.int64), pos_off.cpu().numpy(), values.cpu().numpy()[:total]

    def running_diffs(self, norm_signal, min_seg_len=4):
        """plot_commands.py:225-229: the running-difference trace of a normalised signal window."""
        t = self.torch
        sig = np.ascontiguousarray(norm_signal, dtype=np.float64)
        n = len(sig)
        n_out = max(n - 2 * min_seg_len + 1, 0)
        with t.cuda.device(self.device):
            d_sig = t.from_numpy(sig).to(self.device) if n else t.zeros(1, dtype=t.float64, device=self.device)
            scratch = t.empty(n + 1, dtype=t.float64, device=self.device)
            out = t.empty(max(n_out, 1), dtype=t.float64, device=self.device)
            rc = self.lib.nrq_running_diffs(_ptr(d_sig), n, min_seg_len, _ptr(scratch), _ptr(out), self._stream())
            _lib.check(rc, 'nrq_running_diffs')
        return out.cpu().numpy()[:n_out]

    def event_stats(self, norm_signal, new_segs, compute_sd=True):
        """per-base mean / sd of an already normalised signal (resquiggle.py:62-64), numpy summation
        order, for ONE read.  -> (norm_mean, norm_stdev)."""
        t = self.torch
        sig = np.ascontiguousarray(norm_signal, dtype=np.float64)
        new_segs = np.asarray(new_segs)
        nb = len(new_segs) - 1
        dummy = np.full(max(nb, 1), ord('A'), dtype=np.uint8)
        batch = ReadBatch.from_arrays([np.zeros(1, np.int16)], [0], [np.array([0, len(sig)])], [dummy], [dummy])
        batch.norm_signal = sig
        batch.norm_off = np.array([0, len(sig)], dtype=np.int64)
        db = self.upload(batch)
        pc = ResquiggleParams(compute_sd=compute_sd, exact_stats=True, outlier_thresh=None).to_c()
        with t.cuda.device(self.device):
            sv = t.tensor([0.0, 1.0, float('nan'), float('nan')], dtype=t.float64, device=self.device)
            d_nb = t.tensor([nb], dtype=t.int32, device=self.device)
            d_segs = t.from_numpy(new_segs.astype(np.int32)).to(self.device)
            d_status = t.zeros(1, dtype=t.int32, device=self.device)
            mean = t.empty(max(nb, 1), dtype=t.float64, device=self.device)
            sd = t.empty(max(nb, 1), dtype=t.float64, device=self.device)
            st = t.empty(max(nb, 1), dtype=t.int32, device=self.device)
            ln = t.empty(max(nb, 1), dtype=t.int32, device=self.device)
            rc = self.lib.nrq_event_table(C.byref(db.c), C.byref(pc), _ptr(sv), _ptr(d_nb), _ptr(d_segs),
                                          _ptr(d_status), _ptr(mean), _ptr(sd), _ptr(st), _ptr(ln),
                                          C.c_void_p(0), self._stream())
            _lib.check(rc, 'nrq_event_table')
        return mean.cpu().numpy()[:nb], sd.cpu().numpy()[:nb]


_ENGINE = {}


def get_engine(device=None):
    """process-wide engine per device (created on first use; raises without a GPU / libnrq.so)"""
    import torch
    key = torch.cuda.current_device() if (device is None and torch.cuda.is_available()) else device
    if key not in _ENGINE:
        _ENGINE[key] = ResquiggleEngine(device)
    return _ENGINE[key]
