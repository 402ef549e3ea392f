"""Ragged CSR packing of reads for the device path.

One ``ReadBatch`` holds what ``resquiggle_worker`` dequeues for many reads
(resquiggle.py:425-426) plus their raw DAC signals, concatenated back to back with int64
offsets -- the layout ``include/nrq.h:nrq_batch`` describes.  Host arrays are numpy; the
device copy lives in torch tensors (PyTorch is only the allocator / stream owner here).
"""
import numpy as np


class ReadBatch(object):
    FIELDS = ('raw', 'raw_off', 'read_start', 'starts', 'starts_off',
              'read_align', 'genome_align', 'align_off')

    def __init__(self, raw, raw_off, read_start, starts, starts_off, read_align,
                 genome_align, align_off, scale_values_in=None, norm_signal=None, norm_off=None):
        self.raw = raw
        self.raw_off = raw_off
        self.read_start = read_start
        self.starts = starts
        self.starts_off = starts_off
        self.read_align = read_align
        self.genome_align = genome_align
        self.align_off = align_off
        self.scale_values_in = scale_values_in
        # optional pre-normalised float64 signal (get_indel_groups call surface)
        self.norm_signal = norm_signal
        self.norm_off = norm_off

    @property
    def n_reads(self):
        return len(self.read_start)

    @property
    def total_align(self):
        return int(self.align_off[-1])

    @property
    def n_samples(self):
        """starts_rel_to_read[-1] per read: the samples each read normalises."""
        return self.starts[self.starts_off[1:] - 1].astype(np.int64)

    @classmethod
    def from_arrays(cls, raws, read_starts, starts_list, read_aligns, genome_aligns,
                    scale_values_in=None):
        """raws: int16 arrays; starts_list: integer arrays (Br+1); aligns: bytes / uint8."""
        n = len(raws)

        def offsets(lens):
            off = np.zeros(n + 1, dtype=np.int64)
            np.cumsum(lens, out=off[1:])
            return off

        def as_u8(a):
            return np.frombuffer(a, dtype=np.uint8) if isinstance(a, (bytes, bytearray)) \
                else np.asarray(a, dtype=np.uint8)

        ra = [as_u8(a) for a in read_aligns]
        ga = [as_u8(a) for a in genome_aligns]
        for a, g in zip(ra, ga):
            if len(a) != len(g):
                raise ValueError('alignment rows differ in length')
        raw_off = offsets([len(r) for r in raws])
        starts_off = offsets([len(s) for s in starts_list])
        align_off = offsets([len(a) for a in ra])
        cat = (lambda xs, dt: np.concatenate(xs).astype(dt, copy=False) if n else np.zeros(0, dt))
        for s in starts_list:
            if len(s) and int(s[-1]) > np.iinfo(np.int32).max:
                raise ValueError('read longer than 2^31 samples')
        sv = None
        if scale_values_in is not None:
            sv = np.ascontiguousarray(scale_values_in, dtype=np.float64).reshape(n, 4)
        return cls(cat(raws, np.int16), raw_off,
                   np.asarray(read_starts, dtype=np.int32), cat(starts_list, np.int32),
                   starts_off, cat(ra, np.uint8), cat(ga, np.uint8), align_off, sv)

    @classmethod
    def from_reads(cls, reads, scale_values_in=None):
        """reads: objects with raw, read_start_rel_to_raw, starts, read_align, genome_align
        (nanoraw_b200.synth.SynthRead or the worker's queue items)."""
        return cls.from_arrays(
            [r.raw for r in reads], [r.read_start_rel_to_raw for r in reads],
            [r.starts for r in reads], [r.read_align for r in reads],
            [r.genome_align for r in reads], scale_values_in)

    def input_bytes(self):
        n = sum(getattr(self, f).nbytes for f in self.FIELDS)
        return n + (self.scale_values_in.nbytes if self.scale_values_in is not None else 0)

    def shard(self, rank, world):
        """Contiguous read range for one rank, balanced by total samples."""
        lo, hi = shard_bounds(self.raw_off, rank, world)
        return self.slice(lo, hi)

    def slice(self, lo, hi):
        def cut(data, off):
            return data[off[lo]:off[hi]], off[lo:hi + 1] - off[lo]
        raw, raw_off = cut(self.raw, self.raw_off)
        starts, starts_off = cut(self.starts, self.starts_off)
        ra, align_off = cut(self.read_align, self.align_off)
        ga, _ = cut(self.genome_align, self.align_off)
        sv = None if self.scale_values_in is None else self.scale_values_in[lo:hi]
        return ReadBatch(raw, raw_off, self.read_start[lo:hi], starts, starts_off, ra, ga,
                         align_off, sv)


def shard_bounds(raw_off, rank, world):
    """[lo, hi) read range of `rank` when reads are split into `world` contiguous ranges of
    nearly equal raw-sample totals."""
    n = len(raw_off) - 1
    total = int(raw_off[-1])
    cuts = [int(np.searchsorted(raw_off, total * k / world, side='left')) for k in range(world + 1)]
    cuts[0], cuts[-1] = 0, n
    for k in range(1, world + 1):
        cuts[k] = max(cuts[k], cuts[k - 1])
    return cuts[rank], cuts[rank + 1]


class PinnedBatcher(object):
    """Grow-only page-locked staging for device batches (north_star subsystem (a)).

    Reads are appended straight into pinned ragged CSR buffers -- no per-batch allocation and no
    second copy -- then handed to nrq_resquiggle_batch_host, whose results land in pinned output
    buffers owned here.  Views returned by ``results`` stay valid until the next ``reset``.
    Without CUDA (CPU tests of the packing logic) ordinary numpy memory is used.
    """
    IN = (('raw', np.int16), ('starts', np.int32), ('read_align', np.uint8), ('genome_align', np.uint8))
    OUT = (('norm_mean', np.float64, 1), ('norm_stdev', np.float64, 1), ('ev_start', np.uint32, 1),
           ('ev_length', np.uint32, 1), ('ev_base', np.uint8, 1))

    def __init__(self, max_reads=512, raw_capacity=1 << 26, pinned=True):
        self._pinned = pinned
        self._keep = {}
        self.max_reads = 0
        self._cap = {'raw': 0, 'starts': 0, 'align': 0}
        self._grow_reads(max_reads)
        self._grow('raw', raw_capacity)
        self._grow('starts', raw_capacity // 8)
        self._grow('align', raw_capacity // 8)
        self.reset()

    # ---- storage
    def _alloc(self, name, n, dtype):
        n = max(int(n), 1)
        if self._pinned:
            import torch
            if torch.cuda.is_available():
                t = torch.empty(n * np.dtype(dtype).itemsize, dtype=torch.uint8, pin_memory=True)
                self._keep[name] = t
                return t.numpy().view(dtype)
        return np.empty(n, dtype=dtype)

    def _regrow(self, name, n, dtype, used):
        old = getattr(self, name, None)
        new = self._alloc(name, n, dtype)
        if old is not None and used:
            new[:used] = old[:used]
        setattr(self, name, new)

    def _grow(self, kind, need):
        cap = max(int(need), int(self._cap[kind] * 1.5))
        used = {'raw': self._n_raw, 'starts': self._n_starts, 'align': self._n_align} \
            if hasattr(self, '_n_raw') else {'raw': 0, 'starts': 0, 'align': 0}
        if kind == 'raw':
            self._regrow('raw', cap, np.int16, used['raw'])
        elif kind == 'starts':
            self._regrow('starts', cap, np.int32, used['starts'])
        else:
            self._regrow('read_align', cap, np.uint8, used['align'])
            self._regrow('genome_align', cap, np.uint8, used['align'])
            self._regrow('new_segs', cap + self.max_reads + 1, np.int32, 0)
            for name, dt, _ in self.OUT:
                self._regrow(name, cap, dt, 0)
        self._cap[kind] = cap

    def _grow_reads(self, need):
        n = max(int(need), int(self.max_reads * 1.5))
        used = getattr(self, 'n_reads', 0)
        for name in ('raw_off', 'starts_off', 'align_off'):
            self._regrow(name, n + 1, np.int64, used + 1 if used else 0)
        self._regrow('read_start', n, np.int32, used)
        self._regrow('scale_values_in', 4 * n, np.float64, 4 * used)
        for name in ('status', 'n_bases', 'n_indels', 'region_samples'):
            self._regrow(name, n, np.int32, 0)
        self._regrow('scale_values', 4 * n, np.float64, 0)
        self.max_reads = n
        if self._cap['align']:
            self._regrow('new_segs', self._cap['align'] + n + 1, np.int32, 0)

    # ---- packing
    def reset(self):
        self.n_reads = 0
        self._n_raw = self._n_starts = self._n_align = 0
        self.raw_off[0] = self.starts_off[0] = self.align_off[0] = 0
        self.has_given = False

    @property
    def total_align(self):
        return self._n_align

    def add(self, raw, read_start, starts, read_row, genome_row, given=None):
        """append one read; returns its index in the batch"""
        raw = np.asarray(raw)
        starts = np.asarray(starts)
        ra = np.frombuffer(read_row, dtype=np.uint8) if isinstance(read_row, (bytes, bytearray)) \
            else np.asarray(read_row, dtype=np.uint8)
        ga = np.frombuffer(genome_row, dtype=np.uint8) if isinstance(genome_row, (bytes, bytearray)) \
            else np.asarray(genome_row, dtype=np.uint8)
        if len(ra) != len(ga):
            raise ValueError('alignment rows differ in length')
        if len(starts) and int(starts[-1]) > np.iinfo(np.int32).max:
            raise ValueError('read longer than 2^31 samples')
        r = self.n_reads
        if r + 1 > self.max_reads:
            self._grow_reads(r + 1)
        if self._n_raw + len(raw) > self._cap['raw']:
            self._grow('raw', self._n_raw + len(raw))
        if self._n_starts + len(starts) > self._cap['starts']:
            self._grow('starts', self._n_starts + len(starts))
        if self._n_align + len(ra) > self._cap['align']:
            self._grow('align', self._n_align + len(ra))
        self.raw[self._n_raw:self._n_raw + len(raw)] = raw
        self.starts[self._n_starts:self._n_starts + len(starts)] = starts
        self.read_align[self._n_align:self._n_align + len(ra)] = ra
        self.genome_align[self._n_align:self._n_align + len(ga)] = ga
        self._n_raw += len(raw)
        self._n_starts += len(starts)
        self._n_align += len(ra)
        self.raw_off[r + 1] = self._n_raw
        self.starts_off[r + 1] = self._n_starts
        self.align_off[r + 1] = self._n_align
        self.read_start[r] = read_start
        if given is not None:
            self.scale_values_in[4 * r:4 * r + 4] = (given[0], given[1], np.nan, np.nan) \
                if len(given) == 2 else given
            self.has_given = True
        self.n_reads = r + 1
        return r

    def fill(self, raw_lens, read_starts, starts_list, read_rows, genome_rows, given=None):
        """A whole batch at once, the raw signals left for the caller to write: replaces what the batcher holds by R
        reads with raw_lens[i] samples each and returns raw_off (int64[R + 1]); self.raw[raw_off[i]:raw_off[i + 1]]
        is read i's signal (libnrqio inflates FAST5 signals straight into it).  Rows are bytes; `given`: None or
        R pairs (shift, scale).  The caller has checked what `add` checks per read."""
        R = len(raw_lens)
        self.reset()
        if R > self.max_reads:
            self._grow_reads(R)
        raw_off = np.zeros(R + 1, dtype=np.int64)
        np.cumsum(np.asarray(raw_lens, dtype=np.int64), out=raw_off[1:])
        st_off = np.zeros(R + 1, dtype=np.int64)
        np.cumsum(np.fromiter((len(s) for s in starts_list), dtype=np.int64, count=R), out=st_off[1:])
        al_off = np.zeros(R + 1, dtype=np.int64)
        np.cumsum(np.fromiter((len(s) for s in read_rows), dtype=np.int64, count=R), out=al_off[1:])
        n_raw, n_st, n_al = int(raw_off[R]), int(st_off[R]), int(al_off[R])
        if n_raw > self._cap['raw']:
            self._grow('raw', n_raw)
        if n_st > self._cap['starts']:
            self._grow('starts', n_st)
        if n_al > self._cap['align']:
            self._grow('align', n_al)
        if R:
            self.starts[:n_st] = np.concatenate([np.asarray(s) for s in starts_list])
            self.read_align[:n_al] = np.frombuffer(b''.join(read_rows), dtype=np.uint8)
            self.genome_align[:n_al] = np.frombuffer(b''.join(genome_rows), dtype=np.uint8)
            self.read_start[:R] = np.asarray(read_starts, dtype=np.int64)
        self.raw_off[:R + 1] = raw_off
        self.starts_off[:R + 1] = st_off
        self.align_off[:R + 1] = al_off
        if given is not None and R:
            sv = self.scale_values_in[:4 * R].reshape(R, 4)
            sv[:, :2] = np.asarray(given, dtype=np.float64).reshape(R, 2)
            sv[:, 2:] = np.nan
            self.has_given = True
        self._n_raw, self._n_starts, self._n_align = n_raw, n_st, n_al
        self.n_reads = R
        return raw_off

    def input_bytes(self):
        return 2 * self._n_raw + 4 * self._n_starts + 2 * self._n_align + 28 * self.n_reads
