"""CPU oracle: Python 3 / numpy restatement of nanoraw's per-read re-squiggle path.

TEST INFRASTRUCTURE ONLY -- never imported by ``nanoraw_b200``.  Used by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs.

Parity status: PINNED.  The reference (Python 2, /root/reference) cannot be
imported as is, but ``oracle/run_reference.py`` executes the reference's own
source through a small in-memory Py2->Py3 shim; ``tests/golden/`` holds vectors
produced by that real code and ``tests/test_oracle_vs_reference.py`` checks this
restatement against them (and live against /root/reference when it is present).

Each function cites the reference lines (relative to /root/reference/nanoraw/)
it restates.  Two deliberate additions, both clearly marked:

* ``sort_kind``: the reference visits running differences in
  ``argsort()[::-1]`` order (resquiggle.py:246) -- numpy's default, unstable,
  CPU-dispatch dependent sort.  ``sort_kind='default'`` reproduces that call on
  this host; ``sort_kind='stable'`` (the deterministic contract the CUDA path
  is held to) uses ``argsort(kind='stable')[::-1]``: larger
  value first, larger index first among equal values.
* ``RegionUnsatisfiable``: the reference loops forever when a region already
  spans the whole read and still cannot be satisfied (resquiggle.py:210-215,
  262-273).  The oracle raises instead.
"""
import re
from collections import namedtuple
from time import time

import numpy as np

# record types -- resquiggle.py:29-38, nanoraw_helper.py:18-23
indelStats = namedtuple('indelStats', ('start', 'end', 'diff'))
indelGroupStats = namedtuple('indelGroupStats', ('start', 'end', 'cpts', 'indels'))
scaleValues = namedtuple('scaleValues', ('shift', 'scale', 'lower_lim', 'upper_lim'))
channelInfo = namedtuple(
    'channelInfo', ('offset', 'range', 'digitisation', 'number', 'sampling_rate'))

NORM_TYPES = ('none', 'pA', 'pA_raw', 'median', 'robust_median')  # nanoraw_helper.py:25
_GAP_RUN = re.compile('-+')  # resquiggle.py:49

EVENT_DTYPE = [('norm_mean', '<f8'), ('norm_stdev', '<f8'),
               ('start', '<u4'), ('length', '<u4'), ('base', 'S1')]  # resquiggle.py:69-70

MSG_CPTS_LIMIT = 'Reached maximum number of changepoints for a single indel'  # :235
MSG_TIMEOUT = 'Read took too long to re-segment.'  # :289
MSG_ZERO_LEN = 'New segments include zero length events.'  # :375
MSG_NEG_START = 'New segments start with negative index.'  # :378
MSG_PAST_END = 'New segments end past raw signal values.'  # :381
MSG_SEQ_MISMATCH = 'Aligned sequence does not match number of segments produced.'  # :386
MSG_EVENTS = 'Error computing new events.'  # :77
MSG_UNSATISFIABLE = ('Indel region spans the whole read and still cannot be '
                     're-segmented (the reference does not terminate here).')


class RegionUnsatisfiable(RuntimeError):
    """Raised where resquiggle.py:210-215 / :262-273 would spin forever."""


# --------------------------------------------------------------------------
# a1 / a1'  -- nanoraw_helper.py:184-264
# --------------------------------------------------------------------------
def calc_kmer_fitted_shift_scale(pore_model, events_means, events_kmers):
    """nanoraw_helper.py:184-203 -- weighted least squares of event means on
    the k-mer model means (weights = inverse model variance)."""
    mu = np.array([pore_model['mean'][k] for k in events_kmers])
    w = np.array([pore_model['inv_var'][k] for k in events_kmers])
    mu_w = mu * w
    s_mu_w = mu_w.sum()
    lhs = np.array(((w.sum(), s_mu_w), (s_mu_w, (mu_w * mu).sum())))
    ev_w = events_means * w
    rhs = np.array((ev_w.sum(), (ev_w * mu).sum()))
    shift, scale = np.linalg.solve(lhs, rhs)
    return shift, scale


def normalize_raw_signal(
        all_raw_signal, read_start_rel_to_raw, read_obs_len,
        norm_type=None, channel_info=None, outlier_thresh=None,
        shift=None, scale=None, lower_lim=None, upper_lim=None,
        pore_model=None, event_means=None, event_kmers=None,
        verbatim_clip_loop=False):
    """nanoraw_helper.py:205-264.

    ``verbatim_clip_loop`` keeps the per-sample Python loop of :257-262 (used
    when this oracle is timed as the CPU baseline); the vectorised ``np.where``
    used otherwise yields the same values.
    The reference runs under ``np.seterr(all='raise')`` (resquiggle.py:8).
    """
    if norm_type not in NORM_TYPES and (shift is None or scale is None):
        raise NotImplementedError(
            'Normalization type ' + norm_type + ' is not a valid ' +
            'option and shift or scale parameters were not provided.')
    with np.errstate(all='raise'):
        sig = np.array(all_raw_signal[
            read_start_rel_to_raw:read_start_rel_to_raw + read_obs_len])  # :215-217
        if shift is None or scale is None:
            if norm_type == 'none':  # :219-220
                shift, scale = 0, 1
            elif norm_type in ('pA_raw', 'pA'):  # :221-236
                shift, scale = (-1 * channel_info.offset,
                                channel_info.digitisation / channel_info.range)
                if norm_type == 'pA':
                    fit_shift, fit_scale = calc_kmer_fitted_shift_scale(
                        pore_model, event_means, event_kmers)
                    shift = shift + (fit_shift * scale)
                    scale = scale * fit_scale
            elif norm_type == 'median':  # :240-242
                shift = np.median(sig)
                scale = np.median(np.abs(sig - shift))
            elif norm_type == 'robust_median':  # :243-246 (NameError upstream)
                raise NameError("name 'read_robust_med' is not defined")
        sig = (sig - shift) / scale  # :248
        if outlier_thresh is not None or (
                lower_lim is not None and upper_lim is not None):  # :250-251
            if outlier_thresh is not None:  # :252-256
                med = np.median(sig)
                mad = np.median(np.abs(sig - med))
                lower_lim = med - (mad * outlier_thresh)
                upper_lim = med + (mad * outlier_thresh)
            if verbatim_clip_loop:  # :257-262
                sig = np.array([
                    upper_lim if too_high else (lower_lim if too_low else x)
                    for x, too_high, too_low in zip(
                        sig, sig > upper_lim, sig < lower_lim)])
            else:
                sig = np.where(sig > upper_lim, upper_lim,
                               np.where(sig < lower_lim, lower_lim, sig))
                sig = np.asarray(sig, dtype=np.float64)
    return sig, scaleValues(shift, scale, lower_lim, upper_lim)


# --------------------------------------------------------------------------
# a2 -- resquiggle.py:142-201
# --------------------------------------------------------------------------
def get_all_indels(read_align, genome_align):
    """resquiggle.py:142-201.  ``read_align`` / ``genome_align`` are the two rows
    of alignVals as str.  Returns [indelStats] in alignment order; start/end
    index read bases (= indices into starts_rel_to_read)."""
    n_col = len(read_align)
    runs = sorted(
        [(m.start(), m.end()) for m in _GAP_RUN.finditer(genome_align)] +
        [(m.start(), m.end()) for m in _GAP_RUN.finditer(read_align)] +
        [(0, 0), (n_col, n_col)])  # :146-152
    between = [genome_align[runs[k][1]:runs[k + 1][0]]
               for k in range(len(runs) - 1)]  # :153-156
    inner = runs[1:-1]
    gap_in_read = [read_align[s:e].startswith('-') for s, e in inner]  # :158-159
    run_seq = [genome_align[s:e] if ins else read_align[s:e]
               for ins, (s, e) in zip(gap_in_read, inner)]  # :160-164

    out = []
    read_pos = len(between[0])  # :169
    for k, (seq, ins) in enumerate(zip(run_seq, gap_in_read)):
        before, after = between[k], between[k + 1]
        n = len(seq)
        end = read_pos + 1 if ins else read_pos + n + 1  # :178-179
        diff = n if ins else -n  # :180
        up, down = -1, 0  # :187
        while down < len(after) - 1 and seq[down % n] == after[down]:  # :188-190
            down += 1
        while -up <= len(before) - 1 and seq[(up % n) - n] == before[up]:  # :191-193
            up -= 1
        out.append(indelStats(read_pos + up, end + down, diff))  # :194-195
        if not ins:  # :197-199
            read_pos += n
        read_pos += len(after)
    return out


# --------------------------------------------------------------------------
# a3 / a4 / a5 -- resquiggle.py:203-316
# --------------------------------------------------------------------------
def running_diffs(signal_region, min_seg_len=4):
    """resquiggle.py:237-243 -- sequential float64 cumsum restarted at the
    region start, then |2*cs[i+m] - cs[i] - cs[i+2m]| evaluated as
    ((2*b) - a) - c."""
    cs = np.cumsum(np.insert(signal_region, 0, 0))
    return np.abs((2 * cs[min_seg_len:-min_seg_len]) -
                  cs[:-2 * min_seg_len] - cs[2 * min_seg_len:])


class _Resegmenter(object):
    """State of one get_indel_groups call (the nested functions of
    resquiggle.py:203-276 share align_segs/raw_signal/indel_groups by closure;
    here they are attributes)."""

    def __init__(self, align_segs, signal, min_seg_len, num_cpts_limit,
                 sort_kind, stats):
        self.segs = align_segs
        self.signal = signal
        self.m = min_seg_len
        self.limit = num_cpts_limit
        self.sort_kind = sort_kind
        self.done = []  # finished indelGroupStats ("indel_groups")
        self.stats = stats
        self.last = len(align_segs) - 1

    def _grow(self, start, stop, num_cpts):
        """one step of :212-215 == :263-266; returns None if nothing can grow."""
        if start == 0 and stop == self.last:
            raise RegionUnsatisfiable(MSG_UNSATISFIABLE)
        num_cpts += int(start > 0) + int(stop < self.last)
        return max(0, start - 1), min(self.last, stop + 1), num_cpts

    def extend_group(self, indels):  # :203-216
        start = min(i.start for i in indels)
        stop = max(i.end for i in indels)
        num_cpts = sum(i.diff for i in indels) + stop - start - 1
        while self.segs[stop] - self.segs[start] < (num_cpts + 2) * self.m:
            start, stop, num_cpts = self._grow(start, stop, num_cpts)
        return start, stop, num_cpts

    def _swallow_previous(self, start, stop, num_cpts, indels):  # :220-225 == :267-272
        while len(self.done) > 0 and start <= self.done[-1].end:
            indels = self.done[-1].indels + indels
            del self.done[-1]
            start, stop, num_cpts = self.extend_group(indels)
            if self.stats is not None:
                self.stats['merges'] += 1
        return start, stop, num_cpts, indels

    def extend_and_join(self, indels):  # :217-226
        start, stop, num_cpts = self.extend_group(indels)
        return self._swallow_previous(start, stop, num_cpts, indels)

    def get_cpts(self, start, stop, num_cpts):  # :227-255
        if self.limit is not None and num_cpts > self.limit:
            raise RuntimeError(MSG_CPTS_LIMIT)
        if self.stats is not None:
            self.stats['calls'] += 1
            self.stats['region_sizes'].append(
                int(self.segs[stop] - self.segs[start]))
            self.stats['num_cpts'].append(int(num_cpts))
        rd = running_diffs(
            self.signal[self.segs[start]:self.segs[stop]], self.m)
        if self.sort_kind == 'default':
            order = rd.argsort()[::-1]  # :246 as written
        else:
            order = rd.argsort(kind='stable')[::-1]
        picked = []
        banned = set()
        for pos in order:
            if pos not in banned:
                picked.append(pos)
                banned.update(range(pos - self.m + 1, pos + self.m + 1))
            if len(picked) == num_cpts:
                break
        if len(picked) < num_cpts:
            return None
        return sorted([p + self.m for p in picked])

    def extend_for_cpts(self, start, stop, num_cpts, indels):  # :256-276
        cpts = self.get_cpts(start, stop, num_cpts)
        while cpts is None or cpts[0] == 0:
            if self.stats is not None:
                self.stats['retries'] += 1
            start, stop, num_cpts = self._grow(start, stop, num_cpts)
            start, stop, num_cpts, indels = self._swallow_previous(
                start, stop, num_cpts, indels)
            cpts = self.get_cpts(start, stop, num_cpts)
        return [c + self.segs[start] for c in cpts], start, stop, indels


def get_indel_groups(alignVals, align_segs, raw_signal, min_seg_len, timeout,
                     num_cpts_limit, sort_kind='stable', stats=None):
    """resquiggle.py:139-316.  ``alignVals`` is either the reference's list of
    (read_base, genome_base) pairs or a (read_align, genome_align) str pair."""
    if isinstance(alignVals, tuple):  # (read_align, genome_align) rows
        read_align, genome_align = alignVals
    else:
        read_align = ''.join(rb for rb, _ in alignVals)
        genome_align = ''.join(gb for _, gb in alignVals)
    if timeout is not None:
        t0 = time()
    all_indels = get_all_indels(read_align, genome_align)  # :282
    if len(all_indels) == 0:
        return []
    if stats is not None:
        stats.setdefault('calls', 0)
        stats.setdefault('retries', 0)
        stats.setdefault('merges', 0)
        stats.setdefault('region_sizes', [])
        stats.setdefault('num_cpts', [])
    rs = _Resegmenter(align_segs, raw_signal, min_seg_len, num_cpts_limit,
                      sort_kind, stats)
    cur = [all_indels[0], ]
    for indel in all_indels[1:]:  # :287-304
        if timeout is not None and time() - t0 > timeout:
            raise RuntimeError(MSG_TIMEOUT)
        if max(g.end for g in cur) >= indel.start:
            cur.append(indel)
            continue
        start, stop, num_cpts, cur = rs.extend_and_join(cur)
        cpts, start, stop, cur = rs.extend_for_cpts(start, stop, num_cpts, cur)
        if stop >= indel.start:
            cur.append(indel)
        else:
            rs.done.append(indelGroupStats(start, stop, cpts, cur))
            cur = [indel, ]
    if len(rs.done) == 0 or rs.done[-1].indels[-1] != all_indels[-1]:  # :307-314
        start, stop, num_cpts, cur = rs.extend_and_join(cur)
        cpts, start, stop, cur = rs.extend_for_cpts(start, stop, num_cpts, cur)
        rs.done.append(indelGroupStats(start, stop, cpts, cur))
    return rs.done


# --------------------------------------------------------------------------
# a6 -- resquiggle.py:362-387
# --------------------------------------------------------------------------
def splice_segments(starts_rel_to_read, indel_groups, norm_signal_len,
                    genome_align):
    pieces = []
    prev_stop = 0
    for g_start, g_stop, cpts, _ in indel_groups:  # :364-369
        pieces.append(np.append(starts_rel_to_read[prev_stop:g_start + 1], cpts))
        prev_stop = g_stop
    pieces.append(starts_rel_to_read[prev_stop:])  # :371
    new_segs = np.concatenate(pieces)
    if min(np.diff(new_segs)) < 1:
        raise NotImplementedError(MSG_ZERO_LEN)
    if new_segs[0] < 0:
        raise NotImplementedError(MSG_NEG_START)
    if new_segs[-1] > norm_signal_len:
        raise NotImplementedError(MSG_PAST_END)
    align_seq = genome_align.replace('-', '')  # :384
    if new_segs.shape[0] != len(align_seq) + 1:
        raise ValueError(MSG_SEQ_MISMATCH)
    return new_segs, align_seq


# --------------------------------------------------------------------------
# a7 -- resquiggle.py:60-77
# --------------------------------------------------------------------------
def event_table(norm_signal, new_segs, align_seq, compute_sd=True):
    try:
        with np.errstate(all='raise'):
            stats = [(seg.mean(), seg.std() if compute_sd else np.nan)
                     for seg in np.split(norm_signal, new_segs[1:-1])]  # :62-64
        means = [m for m, _ in stats]
        sds = [s for _, s in stats]
        return np.array(
            list(zip(means, sds, new_segs[:-1], np.diff(new_segs),
                     list(align_seq))), dtype=EVENT_DTYPE)  # :66-70
    except Exception:
        raise NotImplementedError(MSG_EVENTS)


# --------------------------------------------------------------------------
# whole read, without FAST5 I/O -- resquiggle.py:352-396
# --------------------------------------------------------------------------
def resquiggle_read(all_raw_signal, read_start_rel_to_raw, starts_rel_to_read,
                    read_align, genome_align, norm_type='median',
                    outlier_thresh=5, timeout=None, num_cpts_limit=None,
                    compute_sd=True, channel_info=None, pore_model=None,
                    event_means=None, event_kmers=None, min_event_obs=4,
                    sort_kind='stable', verbatim_clip_loop=False, stats=None):
    """Everything resquiggle_read (resquiggle.py:318-401) computes between the
    FAST5 read and the FAST5 write.  Raises the reference's exceptions."""
    starts = np.asarray(starts_rel_to_read)
    norm_signal, scale_values = normalize_raw_signal(
        all_raw_signal, read_start_rel_to_raw, starts[-1], norm_type,
        channel_info, outlier_thresh, pore_model=pore_model,
        event_means=event_means, event_kmers=event_kmers,
        verbatim_clip_loop=verbatim_clip_loop)  # :352-355
    groups = get_indel_groups(
        (read_align, genome_align), starts, norm_signal, min_event_obs,
        timeout, num_cpts_limit, sort_kind=sort_kind, stats=stats)  # :358-360
    new_segs, align_seq = splice_segments(
        starts, groups, norm_signal.shape[0], genome_align)  # :362-387
    events = event_table(norm_signal, new_segs, align_seq, compute_sd)  # :391 -> :60-70
    return {
        'norm_signal': norm_signal,
        'scale_values': scale_values,
        'indel_groups': groups,
        'new_segs': new_segs,
        'align_seq': align_seq,
        'events': events,
    }
