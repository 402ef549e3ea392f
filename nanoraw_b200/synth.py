"""Synthetic R9-style reads for the genome_resquiggle hot path.

The reference ships no data (nanoraw/tests/shell_tests.sh:1-6 points at files
that are not in the repository), so parity tests and ``bench.py`` run on reads
made here.  A read is what ``resquiggle_worker`` receives for one
(file, subgroup) pair (resquiggle.py:425-426) plus the raw DAC signal that
``resquiggle_read`` pulls from the FAST5 file (resquiggle.py:332-333):

* ``raw``: int16 DAC samples, including leading / trailing non-read samples
* ``read_start_rel_to_raw``, ``starts`` (= starts_rel_to_read, int64[Br+1])
* ``read_align`` / ``genome_align``: the two rows of alignVals as ASCII bytes

Signal model: the molecule is the genome segment; each genome base dwells
``dwell_min + Geometric`` samples at the level of its 5-mer, plus Gaussian
noise and rare spikes.  The "basecaller" segmentation is the true one damaged
by the alignment's indels (a deleted base loses its boundary, an inserted base
splits a segment), which is the situation the re-squiggle repairs.
"""
import os
from dataclasses import dataclass, field

import numpy as np

_ACGT = np.frombuffer(b'ACGT', dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b'ACGT-N', b'TGCA-N'):
    _COMP[_a] = _b
_CODE = np.zeros(256, dtype=np.int64)
for _i, _a in enumerate(b'ACGT'):
    _CODE[_a] = _i
GAP = ord('-')


@dataclass
class SynthConfig:
    """Knobs of the generator; defaults are SURVEY.md section 8d's standard read."""
    genome_len: int = 4_600_000
    span_mean: float = 10_000.0
    span_sd: float = 1_000.0
    del_rate: float = 0.04      # genome bases missing from the read ('-' in read row)
    ins_rate: float = 0.04      # extra read bases ('-' in genome row)
    mismatch_rate: float = 0.05
    indel_len_p: float = 0.75   # indel run length ~ Geometric(p)
    dwell_mean: float = 8.9
    dwell_min: int = 2
    level_mean: float = 500.0
    level_sd: float = 60.0
    noise_sd: float = 12.0
    spike_rate: float = 5e-4
    spike_sd: float = 400.0
    lead_samples: tuple = (200, 2000)
    trail_samples: int = 100
    clip_bases: tuple = (0, 0)  # soft-clipped read bases each side (uniform range)
    seed_base: int = 1000


@dataclass
class SynthRead:
    raw: np.ndarray
    read_start_rel_to_raw: int
    starts: np.ndarray
    read_align: bytes
    genome_align: bytes
    strand: str = '+'
    genome_start: int = 0
    chrom: str = 'synth_chr1'
    read_id: str = ''
    extras: dict = field(default_factory=dict)

    @property
    def n_samples(self):
        return int(self.starts[-1])

    @property
    def align_vals(self):
        """alignVals exactly as the reference passes it (resquiggle.py:645-656)."""
        return list(zip(self.read_align.decode(), self.genome_align.decode()))

    def tallies(self):
        """readInfo counts as fix_all_clipped_bases computes them (resquiggle.py:476-485)."""
        r = np.frombuffer(self.read_align, dtype=np.uint8)
        g = np.frombuffer(self.genome_align, dtype=np.uint8)
        n_del = int((r == GAP).sum())
        n_ins = int(((g == GAP) & (r != GAP)).sum())
        both = (r != GAP) & (g != GAP)
        n_match = int((both & (r == g)).sum())
        return n_ins, n_del, n_match, int(both.sum()) - n_match


def make_genome(length=4_600_000, seed=0):
    rng = np.random.default_rng(seed)
    return _ACGT[rng.integers(0, 4, size=length)]


def kmer_levels(cfg, seed=12345):
    rng = np.random.default_rng(seed)
    return rng.standard_normal(1024) * cfg.level_sd + cfg.level_mean


def _revcomp(seq):
    return _COMP[seq[::-1]]


def _events(rng, n, rate, p):
    """positions and run lengths of indel events over n genome bases."""
    k = rng.binomial(n, min(1.0, rate * p))
    pos = rng.integers(1, max(2, n - 1), size=k)
    lens = rng.geometric(p, size=k)
    return pos, lens


def make_read(genome, levels, cfg, index, seed=None):
    """One synthetic read; deterministic in (cfg.seed_base, index)."""
    rng = np.random.default_rng(cfg.seed_base + index if seed is None else seed)
    span = max(64, int(rng.normal(cfg.span_mean, cfg.span_sd)))
    clip_l = int(rng.integers(cfg.clip_bases[0], cfg.clip_bases[1] + 1))
    clip_r = int(rng.integers(cfg.clip_bases[0], cfg.clip_bases[1] + 1))
    total = span + clip_l + clip_r
    total = min(total, len(genome) - 1)
    g0 = int(rng.integers(0, len(genome) - total))
    mol = genome[g0:g0 + total]
    strand = '+' if rng.random() < 0.5 else '-'
    if strand == '-':
        mol = _revcomp(mol)
    core = mol[clip_l:total - clip_r]
    n = len(core)

    # ---- alignment operations over the core genome bases ----
    deleted = np.zeros(n, dtype=bool)
    pos, lens = _events(rng, n, cfg.del_rate, cfg.indel_len_p)
    for p_, l_ in zip(pos, lens):
        deleted[p_:p_ + l_] = True
    deleted[0] = deleted[-1] = False
    ins_len = np.zeros(n, dtype=np.int64)
    pos, lens = _events(rng, n, cfg.ins_rate, cfg.indel_len_p)
    np.add.at(ins_len, pos, lens)
    ins_len[-1] = 0
    # the alignment must end on a match column (resquiggle.py:631-637)
    tail = n - 1
    while tail > 0 and deleted[tail - 1]:
        tail -= 1
    # (deleted[n-1] is False, so the last column is a match unless an insertion
    # follows base n-1, which ins_len[-1]=0 excludes)

    # ---- true dwell per molecule base, basecaller segments ----
    p_geo = 1.0 / (cfg.dwell_mean - cfg.dwell_min + 1.0)
    dwell_all = cfg.dwell_min + rng.geometric(p_geo, size=total) - 1
    dwell = dwell_all[clip_l:total - clip_r]
    owner = np.maximum.accumulate(np.where(~deleted, np.arange(n), 0))
    kept = np.flatnonzero(~deleted)
    seg_total = np.bincount(owner, weights=dwell, minlength=n)[kept].astype(np.int64)
    # insertions attach to the owning read base; keep every piece >= 1 sample
    ins_owner = np.bincount(owner, weights=ins_len, minlength=n)[kept].astype(np.int64)
    too_many = ins_owner > seg_total - 1
    if too_many.any():
        for o in kept[too_many]:
            members = np.flatnonzero(owner == o)
            budget = int(dwell[members].sum()) - 1
            for k in members:
                take = min(int(ins_len[k]), budget)
                ins_len[k] = take
                budget -= take
        ins_owner = np.bincount(owner, weights=ins_len, minlength=n)[kept].astype(np.int64)
    seg_lens = []
    for t_o, parts in zip(seg_total.tolist(), (ins_owner + 1).tolist()):
        if parts == 1:
            seg_lens.append(t_o)
        else:
            cuts = np.sort(rng.choice(np.arange(1, t_o), size=parts - 1, replace=False))
            seg_lens.extend(np.diff(np.concatenate(([0], cuts, [t_o]))).tolist())
    starts = np.concatenate(([0], np.cumsum(np.asarray(seg_lens, dtype=np.int64))))

    # ---- alignment rows ----
    counts = 1 + ins_len
    col_base = np.repeat(np.arange(n), counts)
    first = np.ones(len(col_base), dtype=bool)
    first[1:] = col_base[1:] != col_base[:-1]
    read_base = core.copy()
    mm = rng.random(n) < cfg.mismatch_rate
    mm[0] = mm[-1] = False
    read_base[mm] = _ACGT[(_CODE[core[mm]] + rng.integers(1, 4, size=int(mm.sum()))) % 4]
    read_base[deleted] = GAP
    genome_align = np.where(first, core[col_base], GAP).astype(np.uint8)
    ins_bases = _ACGT[rng.integers(0, 4, size=len(col_base))]
    read_align = np.where(first, read_base[col_base], ins_bases).astype(np.uint8)

    # ---- raw DAC signal over the whole molecule (clipped flanks included) ----
    codes = _CODE[mol]
    padded = np.concatenate(([codes[0]] * 2, codes, [codes[-1]] * 2))
    kmer = (padded[:-4] * 256 + padded[1:-3] * 64 + padded[2:-2] * 16 +
            padded[3:-1] * 4 + padded[4:])
    sig = np.repeat(levels[kmer], dwell_all)
    sig = sig + rng.standard_normal(len(sig)) * cfg.noise_sd
    n_lead = int(rng.integers(cfg.lead_samples[0], cfg.lead_samples[1] + 1))
    lead = cfg.level_mean + 150.0 + rng.standard_normal(n_lead) * 30.0
    trail = cfg.level_mean + 150.0 + rng.standard_normal(cfg.trail_samples) * 30.0
    full = np.concatenate((lead, sig, trail))
    spikes = rng.random(len(full)) < cfg.spike_rate
    full[spikes] += rng.standard_normal(int(spikes.sum())) * cfg.spike_sd
    raw = np.clip(np.rint(full), -32768, 32767).astype(np.int16)
    r0 = n_lead + int(dwell_all[:clip_l].sum())

    assert int(starts[-1]) == int(dwell.sum())
    assert len(starts) - 1 == int((read_align != GAP).sum())
    genome_pos = g0 + (clip_l if strand == '+' else clip_r)
    return SynthRead(
        raw=raw, read_start_rel_to_raw=r0, starts=starts,
        read_align=read_align.tobytes(), genome_align=genome_align.tobytes(),
        strand=strand, genome_start=genome_pos,
        read_id='synth_%d_%d' % (cfg.seed_base, index),
        extras={'clip_start': clip_l, 'clip_end': clip_r})


def make_reads(n_reads, cfg=None, genome=None, levels=None, start_index=0):
    cfg = cfg or SynthConfig()
    if genome is None:
        genome = make_genome(cfg.genome_len)
    if levels is None:
        levels = kmer_levels(cfg)
    return [make_read(genome, levels, cfg, start_index + i) for i in range(n_reads)]


def _worker(args):
    cfg, lo, hi = args
    genome = make_genome(cfg.genome_len)
    levels = kmer_levels(cfg)
    return [make_read(genome, levels, cfg, i) for i in range(lo, hi)]


def make_reads_parallel(n_reads, cfg=None, procs=None, start_index=0):
    """Same reads as make_reads, generated over a process pool."""
    import multiprocessing as mp
    import os
    cfg = cfg or SynthConfig()
    procs = procs or min(os.cpu_count() or 1, 32)
    if procs <= 1 or n_reads < 4 * procs:
        return make_reads(n_reads, cfg, start_index=start_index)
    step = -(-n_reads // (procs * 4))
    jobs = [(cfg, start_index + lo, start_index + min(n_reads, lo + step))
            for lo in range(0, n_reads, step)]
    with mp.get_context('fork').Pool(procs) as pool:
        chunks = pool.map(_worker, jobs)
    return [r for c in chunks for r in c]


# named workloads of BASELINE.json:configs
def config_standard(seed_base=1000):
    return SynthConfig(seed_base=seed_base)


def config_short(seed_base=1000):
    """~4.5 kb / ~40k-sample reads (BASELINE.json's literal '~40k samples')."""
    return SynthConfig(span_mean=4500.0, span_sd=450.0, seed_base=seed_base)


def config_long(seed_base=4000):
    return SynthConfig(span_mean=100_000.0, span_sd=10_000.0, seed_base=seed_base)


def config_indel_dense(seed_base=5000):
    return SynthConfig(del_rate=0.12, ins_rate=0.12, clip_bases=(50, 500),
                       seed_base=seed_base)


# ------------------------------------------------------------------------------------------
# Front-end inputs: an Albacore-style basecall Events table and a
# BWA-style SAM record that reproduce a SynthRead when parsed.
BASECALL_EVENT_DTYPE = [('mean', '<f8'), ('start', '<f8'), ('stdv', '<f8'), ('length', '<f8'),
                        ('model_state', 'S5'), ('move', '<i4')]


def to_sam_record(read, soft_clip=(0, 0), hard_clip=(0, 0), mapq=60, seed=0):
    """SAM fields (dict keyed like resquiggle.py:45-47) whose parse gives back read.read_align /
    read.genome_align.  Clips are given in read orientation (start, end)."""
    rng = np.random.default_rng(seed)
    r = np.frombuffer(read.read_align, dtype=np.uint8)
    g = np.frombuffer(read.genome_align, dtype=np.uint8)
    kind = np.where(r == GAP, ord('D'), np.where(g == GAP, ord('I'), ord('M'))).astype(np.uint8)
    cuts = np.flatnonzero(np.diff(kind)) + 1
    bounds = np.concatenate(([0], cuts, [len(kind)]))
    ops = [(int(b - a), chr(kind[a])) for a, b in zip(bounds[:-1], bounds[1:])]
    seq = r[r != GAP]
    lead = _ACGT[rng.integers(0, 4, size=soft_clip[0])]
    tail = _ACGT[rng.integers(0, 4, size=soft_clip[1])]
    seq = np.concatenate((lead, seq, tail)).astype(np.uint8)
    if soft_clip[0]:
        ops.insert(0, (soft_clip[0], 'S'))
    if soft_clip[1]:
        ops.append((soft_clip[1], 'S'))
    if hard_clip[0]:
        ops.insert(0, (hard_clip[0], 'H'))
    if hard_clip[1]:
        ops.append((hard_clip[1], 'H'))
    flag = 0
    if read.strand == '-':  # SAM stores the forward-strand view
        ops = ops[::-1]
        seq = _revcomp(seq)
        flag = 16
    cigar = ''.join('%d%s' % op for op in ops)
    return {'qName': read.read_id or 'read', 'flag': str(flag), 'rName': read.chrom,
            'pos': str(read.genome_start + 1), 'mapq': str(mapq), 'cigar': cigar, 'rNext': '*',
            'pNext': '0', 'tLen': '0', 'seq': seq.tobytes().decode(), 'qual': '*'}


def to_basecall_events(read, stay_rate=0.1, lead_stays=0, trail_stays=0, sampling_rate=4000.0,
                       clip_bases=(0, 0), seed=0):
    """Albacore (> 1.0) style Events for the read's basecalls: one moving event per read base,
    plus `stay` events (move == 0) that split a base's samples, as the basecaller emits them.
    clip_bases adds extra called bases before / after (to be soft-clipped by the mapper).
    Returns (events, start_time_offset_samples)."""
    rng = np.random.default_rng(seed)
    r = np.frombuffer(read.read_align, dtype=np.uint8)
    bases = r[r != GAP]
    seg = np.diff(np.asarray(read.starts, dtype=np.int64))
    if clip_bases[0] or clip_bases[1]:
        pre = _ACGT[rng.integers(0, 4, size=clip_bases[0])]
        post = _ACGT[rng.integers(0, 4, size=clip_bases[1])]
        bases = np.concatenate((pre, bases, post)).astype(np.uint8)
        seg = np.concatenate((rng.integers(3, 15, size=clip_bases[0]), seg,
                              rng.integers(3, 15, size=clip_bases[1])))
    lengths, moves, states = [], [], []
    padded = np.concatenate(([bases[0]] * 2, bases, [bases[-1]] * 2)).astype(np.uint8)

    def kmer(i):
        return padded[i:i + 5].tobytes()

    for _ in range(lead_stays):
        lengths.append(int(rng.integers(2, 9)))
        moves.append(0)
        states.append(kmer(0))
    for i, n in enumerate(seg.tolist()):
        parts = [n]
        if n >= 4 and rng.random() < stay_rate:
            cut = int(rng.integers(1, n))
            parts = [cut, n - cut]
        for j, p in enumerate(parts):
            lengths.append(p)
            moves.append(1 if j == 0 else 0)
            states.append(kmer(i))
    for _ in range(trail_stays):
        lengths.append(int(rng.integers(2, 9)))
        moves.append(0)
        states.append(kmer(len(bases) - 1))
    lengths = np.asarray(lengths, dtype=np.int64)
    ev = np.zeros(len(lengths), dtype=BASECALL_EVENT_DTYPE)
    begin = np.concatenate(([0], np.cumsum(lengths)[:-1]))
    ev['start'] = begin / sampling_rate
    ev['length'] = lengths / sampling_rate
    ev['mean'] = 90.0 + rng.standard_normal(len(lengths)) * 10.0
    ev['stdv'] = 1.5
    ev['model_state'] = states
    ev['move'] = moves
    ev['move'][0] = 1
    return ev


def write_basecalled_fast5_job(job):
    """(file name, SynthRead) -> one basecalled single-read FAST5 file on disk, in the state prep_fast5 leaves it
    (resquiggle.py:942-988); returns its size.  Module-level so that I/O processes can run it."""
    fn, read = job
    from . import fast5_io
    f = fast5_io.open_fast5(fn, 'w')
    rd = f.create_group('Raw/Reads/Read_1')
    rd.create_dataset('Signal', data=read.raw, compression='gzip')
    rd.attrs['read_id'] = read.read_id or os.path.basename(fn)
    rd.attrs['start_time'] = np.uint64(0)
    ch = f.create_group('UniqueGlobalKey/channel_id')
    for k, v in (('offset', 10.0), ('range', 1400.0), ('digitisation', 8192.0), ('sampling_rate', 4000.0)):
        ch.attrs[k] = np.float64(v)
    ch.attrs['channel_number'] = '11'
    f.create_group('Analyses/Basecall_1D_000').attrs['version'] = '1.1.0'
    f.create_dataset('Analyses/Basecall_1D_000/BaseCalled_template/Events',
                     data=to_basecall_events(read, seed=1), compression='gzip')
    grp = f.create_group('Analyses/RawGenomeCorrected_000')  # what prep_fast5 leaves behind (resquiggle.py:942-988)
    grp.attrs['nanoraw_version'] = '0.5'
    grp.attrs['basecall_group'] = 'Basecall_1D_000'
    f.close()
    return os.path.getsize(fn) if os.path.exists(fn) else 0  # (mem:// names have no size)


def drop_corrected_job(job):
    """remove the corrected subgroup again (bench set-up between passes)"""
    from . import fast5_io
    f = fast5_io.open_fast5(job[0], 'r+')
    del f['/Analyses/RawGenomeCorrected_000/BaseCalled_template']
    f.close()
