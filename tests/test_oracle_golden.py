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


def test_default_sort_tie_attribution():
    """Where the reference's own (unstable) argsort order picked different
    change-points on the generating host, the difference must be explained by
    exact ties in the running differences."""
    n_diff = 0
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
            rd = orc.running_diffs(sig[starts[g.start]:starts[g.end]])
            # a differing pick means some value is duplicated in this region
            assert len(np.unique(rd)) < len(rd), (case.name, g.start, g.end)
    assert n_diff > 0
