"""Host mirror of the parts of nanoraw/nanoraw_helper.py the genome_resquiggle path uses.

Same names, argument meaning and error behaviour as the reference (file:line cited per
function); the signal arithmetic runs on the GPU through libnrq (include/nrq.h) -- there is
no numpy fallback for it.
"""
from collections import namedtuple

import numpy as np

from . import _lib
from .errors import exception_for

NANORAW_VERSION = '0.5'  # nanoraw_helper.py:11

channelInfo = namedtuple(
    'channelInfo', ('offset', 'range', 'digitisation', 'number', 'sampling_rate'))  # :18-20
scaleValues = namedtuple('scaleValues', ('shift', 'scale', 'lower_lim', 'upper_lim'))  # :21-23

NORM_TYPES = ('none', 'pA', 'pA_raw', 'median', 'robust_median')  # :25

COMP_BASES = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A', '-': '-', 'N': 'N'}  # :39


def comp_base(base):  # :40-45
    return COMP_BASES.get(base, 'N')


def rev_comp(seq):  # :46-47
    return ''.join(comp_base(b) for b in seq[::-1])


def get_channel_info(fast5_data):
    """nanoraw_helper.py:154-166"""
    try:
        info = fast5_data['UniqueGlobalKey/channel_id'].attrs
    except Exception:
        raise RuntimeError("No channel_id group in HDF5 file. " +
                           "Probably mux scan HDF5 file.")
    return channelInfo(
        info['offset'], info['range'], info['digitisation'], info['channel_number'],
        np.asarray(info['sampling_rate']).astype('int_'))


def parse_pore_model(pore_model_fn):
    """nanoraw_helper.py:168-182 -- k-mer level means and inverse variances."""
    pore_model = {'mean': {}, 'inv_var': {}}
    with open(pore_model_fn) as fp:
        for line in fp:
            if line.startswith('#'):
                continue
            try:
                kmer, lev_mean, lev_stdev = line.split()[:3]
                lev_mean, lev_stdev = float(lev_mean), float(lev_stdev)
            except ValueError:
                continue  # header or other non-kmer field
            pore_model['mean'][kmer] = lev_mean
            pore_model['inv_var'][kmer] = 1 / (lev_stdev * lev_stdev)
    return pore_model


def calc_kmer_fitted_shift_scale(pore_model, events_means, events_kmers):
    """nanoraw_helper.py:184-203 -- 2x2 weighted least-squares fit per read.  Host side: it
    consumes the basecaller's per-event table (tens of thousands of values), produces the two
    scalars the device normalisation takes as input, and is not part of the per-sample path."""
    def as_str(k):
        return k.decode() if isinstance(k, bytes) else k
    mu = np.array([pore_model['mean'][as_str(k)] for k in events_kmers])
    w = np.array([pore_model['inv_var'][as_str(k)] for k in events_kmers])
    mu_w = mu * w
    s_mu_w = mu_w.sum()
    lhs = np.array(((w.sum(), s_mu_w), (s_mu_w, (mu_w * mu).sum())))
    ev_w = events_means * w
    rhs = np.array((ev_w.sum(), (ev_w * mu).sum()))
    shift, scale = np.linalg.solve(lhs, rhs)
    return shift, scale


def given_shift_scale(norm_type, channel_info=None, pore_model=None, event_means=None,
                      event_kmers=None):
    """(shift, scale) for the normalisation types whose constants do not come from the signal
    (nanoraw_helper.py:219-236); None for 'median'."""
    if norm_type == 'none':
        return 0, 1
    if norm_type in ('pA_raw', 'pA'):
        shift, scale = (-1 * channel_info.offset,
                        channel_info.digitisation / channel_info.range)
        if norm_type == 'pA':
            fit_shift, fit_scale = calc_kmer_fitted_shift_scale(
                pore_model, event_means, event_kmers)
            shift = shift + (fit_shift * scale)
            scale = scale * fit_scale
        return shift, scale
    return None


def normalize_raw_signal(
        all_raw_signal, read_start_rel_to_raw, read_obs_len,
        norm_type=None, channel_info=None, outlier_thresh=None,
        shift=None, scale=None, lower_lim=None, upper_lim=None,
        pore_model=None, event_means=None, event_kmers=None):
    """nanoraw_helper.py:205-264 on the GPU: (float64 normalised signal, scaleValues).

    The order statistics come from nrq_norm_stats, the signal from nrq_normalize_signal;
    scale == 0 raises FloatingPointError with numpy's message, as the reference does under
    np.seterr(all='raise') (resquiggle.py:8)."""
    from .batch import ReadBatch
    from .engine import ResquiggleParams, get_engine
    if norm_type not in NORM_TYPES and (shift is None or scale is None):
        raise NotImplementedError(
            'Normalization type ' + str(norm_type) + ' is not a valid ' +
            'option and shift or scale parameters were not provided.')
    if norm_type == 'robust_median' and (shift is None or scale is None):
        raise NameError("name 'read_robust_med' is not defined")  # nanoraw_helper.py:246
    raw = np.ascontiguousarray(all_raw_signal, dtype=np.int16)
    n_obs = int(read_obs_len)
    sv_in = None
    kind = 'median'
    if shift is None or scale is None:
        given = given_shift_scale(norm_type, channel_info, pore_model, event_means, event_kmers)
        if given is not None:
            shift, scale = given
    if shift is not None and scale is not None:
        kind = 'given'
    given_limits = outlier_thresh is None and lower_lim is not None and upper_lim is not None
    if kind == 'given' or given_limits:
        sv_in = np.array([[np.nan if shift is None else float(shift),
                           np.nan if scale is None else float(scale),
                           float(lower_lim) if given_limits else np.nan,
                           float(upper_lim) if given_limits else np.nan]])
    batch = ReadBatch.from_arrays(
        [raw], [int(read_start_rel_to_raw)], [np.array([0, n_obs], dtype=np.int64)],
        [b'A'], [b'A'], sv_in)
    engine = get_engine()
    params = ResquiggleParams(norm_type=kind, outlier_thresh=outlier_thresh,
                              given_limits=given_limits)
    status, sv, sigs = engine.normalize_signal(engine.upload(batch), params)
    st = int(status[0])
    if st in (_lib.ST_FPE_DIVIDE, _lib.ST_FPE_INVALID, _lib.ST_BAD_INPUT):
        raise exception_for(st)
    out_shift = shift if kind == 'given' else sv[0][0]
    out_scale = scale if kind == 'given' else sv[0][1]
    lims = [None if np.isnan(v) else v for v in sv[0][2:]]
    if given_limits:
        lims = [lower_lim, upper_lim]
    return sigs[0], scaleValues(out_shift, out_scale, lims[0], lims[1])


def parse_fasta(fasta_fn):
    """nanoraw_helper.py:266-296 -- {record id: sequence}."""
    records = {}
    cur_id, chunks = None, []
    with open(fasta_fn) as fp:
        for line in fp:
            if line.startswith('>'):
                if cur_id is not None and chunks:
                    records[cur_id] = ''.join(chunks)
                chunks = []
                cur_id = line.replace('>', '').strip().split()[0]
            else:
                chunks.append(line.strip())
    if cur_id is not None and chunks:
        records[cur_id] = ''.join(chunks)
    return records


# ------------------------------------------------------------------------------------------
# f-3: corrected-read index and genome-anchored aggregation (nanoraw_helper.py:15-17, 110-152, 298-444).
# The per-position arithmetic runs on the GPU (nrq_genome_aggregate); reading the Events back from
# the FAST5 files stays on the host, as in the reference.
readData = namedtuple('readData', (
    'start', 'end', 'segs', 'read_start_rel_to_raw', 'strand', 'fn', 'corr_group'))


def parse_fast5s(files, corrected_group, basecall_subgroups):
    """nanoraw_helper.py:110-152 -> {(chrom, strand): [readData]} for every corrected read found."""
    import sys
    from collections import defaultdict
    from . import fast5_io
    raw_read_coverage = defaultdict(list)
    for read_fn in files:
        for subgroup in basecall_subgroups:
            try:
                read_data = fast5_io.open_fast5(read_fn, 'r')
            except IOError:
                continue  # probably a truncated file
            slot = '/'.join(('/Analyses', corrected_group, subgroup))
            if slot not in read_data:
                continue
            corr = read_data[slot]
            try:
                align = dict(corr['Alignment'].attrs.items())
                read_start_rel_to_raw = corr['Events'].attrs['read_start_rel_to_raw']
                events = fast5_io.read_dataset(corr['Events'])
                events_end = events[-1]['start'] + events[-1]['length']
                segs = np.concatenate([events['start'], [events_end, ]])
            except Exception:
                sys.stderr.write('********** WARNING: Corrupted corrected events in ' +
                                 read_fn + '. ********\n')
                continue
            raw_read_coverage[(align['mapped_chrom'], align['mapped_strand'])].append(readData(
                align['mapped_start'], align['mapped_start'] + len(segs) - 1, segs,
                read_start_rel_to_raw, align['mapped_strand'], read_fn,
                corrected_group + '/' + subgroup))
            read_data.close()
    return raw_read_coverage


def _read_event_columns(r_data):
    from . import fast5_io
    try:
        read_data = fast5_io.open_fast5(r_data.fn, 'r')
    except IOError:
        return None
    slot = '/'.join(('/Analyses', r_data.corr_group, 'Events'))
    if slot not in read_data:
        return None
    events = fast5_io.read_dataset(read_data[slot])
    return events['norm_mean'], events['norm_stdev'], events['length']


def get_base_stats(raw_read_coverage, chrm_sizes):
    """One pass over the corrected reads -> (coverage, mean of norm_mean, mean of norm_stdev, mean of length),
    each a {(chrom, strand): array of chromosome length}.  Additive helper: the three reference functions
    below re-open every FAST5 file once each; this reads them once and aggregates on the GPU."""
    from .engine import get_engine
    engine = get_engine()
    out = ({}, {}, {}, {})
    for chrm in chrm_sizes.keys():
        for strand in ('+', '-'):
            reads = raw_read_coverage.get((chrm, strand), [])
            starts, cols = [], []
            for r_data in reads:
                c = _read_event_columns(r_data)
                if c is None:
                    continue
                starts.append(r_data.start)
                cols.append(c)
            res = engine.genome_aggregate(
                starts, [len(c[0]) for c in cols], [strand == '-'] * len(cols),
                [c[0] for c in cols], [c[1] for c in cols], [c[2] for c in cols], chrm_sizes[chrm])
            for d, arr in zip(out, res):
                d[(chrm, strand)] = arr
    return out


def get_reads_events(chrm_strand_reads, rev_strand):
    """nanoraw_helper.py:446-484: {genomic position: event means of the reads that cover it} for reads of ONE
    chromosome and strand; None when no read has an event table.  The lists are put together on the GPU
    (nrq_genome_event_lists).  The reference orders the means of one position by an unstable argsort of the
    positions, i.e. in no specified order; here they come in the order of the reads' mapped starts."""
    from .engine import get_engine
    starts, means = [], []
    for r_data in chrm_strand_reads:
        c = _read_event_columns(r_data)
        if c is None:
            continue
        starts.append(r_data.start)
        means.append(c[0])
    if not starts:
        return None
    track_len = max(s + len(m) for s, m in zip(starts, means))
    cov, off, values = get_engine().genome_event_lists(
        starts, [len(m) for m in means], [bool(rev_strand)] * len(means), means, track_len)
    return dict((int(p), values[off[p]:off[p + 1]]) for p in np.flatnonzero(cov > 0))


def get_coverage(raw_read_coverage):
    """nanoraw_helper.py:298-308: coverage over [start, end) of every read, up to the largest end."""
    from .engine import get_engine
    engine = get_engine()
    read_coverage = {}
    for (chrm, strand), reads in raw_read_coverage.items():
        max_end = max(r.end for r in reads)
        lens = [r.end - r.start for r in reads]
        zeros = [np.zeros(n) for n in lens]
        cov = engine.genome_aggregate([r.start for r in reads], lens, [False] * len(reads), zeros, zeros,
                                      [np.zeros(n, dtype=np.uint32) for n in lens], max_end)[0]
        read_coverage[(chrm, strand)] = cov.astype(np.int_)
    return read_coverage


def get_base_means(raw_read_coverage, chrm_sizes):
    """nanoraw_helper.py:342-361 (NaN where no read covers the base)"""
    return get_base_stats(raw_read_coverage, chrm_sizes)[1]


def get_base_sds(raw_read_coverage, chrm_sizes):
    """nanoraw_helper.py:388-402"""
    return get_base_stats(raw_read_coverage, chrm_sizes)[2]


def get_base_lengths(raw_read_coverage, chrm_sizes):
    """nanoraw_helper.py:430-444"""
    return get_base_stats(raw_read_coverage, chrm_sizes)[3]
