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
