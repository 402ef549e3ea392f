"""The oracle restatement against vectors computed by the REAL reference
(tests/golden/hotpath_v1.npz, see tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import resquiggle_oracle as orc
from golden_util import load_golden

INFO, CASES = load_golden()


def _channel():
    c = INFO['channel']
    return orc.channelInfo(c['offset'], c['range'], c['digitisation'],
                           c['channel_number'], int(c['sampling_rate']))


def _run(case, sort_kind):
    r = case.read
    return orc.resquiggle_read(
        r.raw, r.read_start_rel_to_raw, r.starts, r.read_align.decode(),
        r.genome_align.decode(), channel_info=_channel(), sort_kind=sort_kind,
        **case.kwargs())


@pytest.mark.parametrize('case', CASES, ids=[c.name for c in CASES])
def test_oracle_matches_reference_stable(case):
    if case.error:
        with pytest.raises(Exception) as ei:
            _run(case, 'stable')
        assert '%s: %s' % (type(ei.value).__name__, ei.value) == case.error
        return
    out = _run(case, 'stable')
    a = case.a
    sv = np.array([np.nan if v is None else float(v) for v in out['scale_values']])
    assert np.array_equal(sv, a['scale_values'], equal_nan=True)
    if 'norm_signal' in a:
        assert np.array_equal(out['norm_signal'], a['norm_signal'])
    indels = np.array([tuple(i) for g in out['indel_groups'] for i in g.indels],
                      dtype=np.int64).reshape(-1, 3)
    assert np.array_equal(indels, a['indels'])
    groups = np.array([(g.start, g.end, len(g.cpts)) for g in out['indel_groups']],
                      dtype=np.int64).reshape(-1, 3)
    assert np.array_equal(groups, a['groups'])
    cpts = np.array([c for g in out['indel_groups'] for c in g.cpts], dtype=np.int64)
    assert np.array_equal(cpts, a['cpts'])
    if 'ev_mean' in a:  # the full route (resquiggle_read + event table) succeeded
        ev = out['events']
        assert np.array_equal(ev['start'], a['ev_start'])
        assert np.array_equal(ev['length'], a['ev_length'])
        assert np.array_equal(ev['base'], a['ev_base'])
        assert np.array_equal(ev['norm_mean'], a['ev_mean'])
        assert np.array_equal(ev['norm_stdev'], a['ev_sd'], equal_nan=True)
        assert np.array_equal(a['read_segments'], case.read.starts)


def _reachable_by_tie_breaks(rd, K, target, m=4):
    """can the greedy selection (resquiggle.py:244-254) end on exactly `target` if, whenever several candidates
    share the largest running difference EXACTLY, it may take any of them?  (That is all an unstable argsort can
    change: positions are visited in descending value, equal values in arbitrary order.)"""
    n = len(rd)
    target = frozenset(target)

    def go(black, picked):
        if len(picked) == K:
            return frozenset(picked) == target
        free = [i for i in range(n) if i not in black]
        if not free:
            return False
        top = max(rd[i] for i in free)
        ties = [i for i in free if rd[i] == top]
        for i in ties:
            if i not in target:
                continue  # a pick outside the target can never be undone
            if go(black | set(range(i - m + 1, i + m + 1)), picked + [i]):
                return True
        return False

    return go(frozenset(), [])


def test_default_sort_tie_attribution():
    """Where the reference's own (unstable) argsort order picked different change-points on the generating host,
    the difference must be explained by EXACT ties in the running differences: the default
    result has to be reachable from the same running differences by breaking ties between exactly equal maxima
    differently, and the first pick at which the two selections part must be such a tie."""
    n_diff = n_groups = 0
    for case in CASES:
        if case.error or not case.meta['default_differs']:
            continue
        n_diff += 1
        out = _run(case, 'stable')
        sig = out['norm_signal']
        starts = case.read.starts
        d_cpts = case.a['default_cpts']
        pos = 0
        for g in out['indel_groups']:
            mine = list(g.cpts)
            theirs = list(d_cpts[pos:pos + len(mine)])
            pos += len(mine)
            if mine == theirs:
                continue
            n_groups += 1
            t0 = int(starts[g.start])
            rd = orc.running_diffs(sig[t0:starts[g.end]])
            K = len(mine)
            want = [c - t0 - 4 for c in theirs]  # change-point -> running-difference index
            assert _reachable_by_tie_breaks(rd, K, want), (case.name, g.start, g.end)
            # the stable selection, pick by pick: the first pick that is not in the default set ties exactly with a
            # candidate that is
            black, default_set = set(), set(want)
            order = sorted(range(len(rd)), key=lambda i: (rd[i], i), reverse=True)
            for _ in range(K):
                p = next(i for i in order if i not in black)
                if p not in default_set:
                    rivals = [i for i in default_set if i not in black and rd[i] == rd[p]]
                    assert rivals, (case.name, g.start, g.end, p)
                    break
                black.update(range(p - 3, p + 5))
    assert n_diff > 0 and n_groups > 0
