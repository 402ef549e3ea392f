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
