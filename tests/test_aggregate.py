"""f-3: genome-anchored aggregation (nanoraw_helper.py:110-152, 298-444) against vectors the REAL reference
computed (tests/golden/aggregate_v1.npz, tests/golden/make_aggregate_golden.py)."""
import json
import os

import numpy as np
import pytest

from nanoraw_b200 import fast5_io
from oracle import aggregate_oracle as agg

HERE = os.path.dirname(os.path.abspath(__file__))
Z = np.load(os.path.join(HERE, 'golden', 'aggregate_v1.npz'))
META = json.loads(Z['meta'].tobytes().decode())
TRACKS = [tuple(t) for t in META['tracks']]


def reads_of(track):
    return [m for m in META['reads'] if m['track'] == track]


@pytest.mark.parametrize('k', range(len(TRACKS)), ids=['%s%s' % t for t in TRACKS])
def test_oracle_matches_reference(k):
    reads = reads_of(k)
    rev = TRACKS[k][1] == '-'
    assert np.array_equal(agg.get_coverage([(m['start'], m['end']) for m in reads]), Z['t%d/coverage' % k])
    L = META['genome_len']
    for col, name in (('norm_mean', 'means'), ('norm_stdev', 'sds'), ('length', 'lengths')):
        key = 't%d/%s' % (k, name)
        if key not in Z.files:
            continue
        got = agg.mean_over_reads([(m['start'], Z[m['tag'] + 'events'][col]) for m in reads], L, rev)
        assert np.array_equal(got, Z[key], equal_nan=True)


def _write_corrected_files():
    """rebuild the corrected groups the reference wrote, from the stored event tables"""
    files = []
    for i, m in enumerate(META['reads']):
        name = 'mem://agg_test/%d' % i
        f = fast5_io.MemFile(name, 'w')
        sub = f.create_group('Analyses/RawGenomeCorrected_000/BaseCalled_template')
        al = sub.create_group('Alignment')
        al.attrs['mapped_start'] = m['start']
        al.attrs['mapped_strand'] = m['strand']
        al.attrs['mapped_chrom'] = m['chrom']
        ev = sub.create_dataset('Events', data=Z[m['tag'] + 'events'])
        ev.attrs['read_start_rel_to_raw'] = m['read_start_rel_to_raw']
        f.close()
        files.append(name)
    fast5_io.MemFile('mem://agg_test/uncorrected', 'w').close()
    return files + ['mem://agg_test/uncorrected', 'mem://agg_test/missing']


def test_parse_fast5s_indexes_corrected_reads():
    from nanoraw_b200 import nanoraw_helper as nh
    rrc = nh.parse_fast5s(_write_corrected_files(), 'RawGenomeCorrected_000', ['BaseCalled_template'])
    assert sorted(rrc) == TRACKS
    for k, key in enumerate(TRACKS):
        want = reads_of(k)
        assert [(r.start, r.end) for r in rrc[key]] == [(m['start'], m['end']) for m in want]
        for r, m in zip(rrc[key], want):
            ev = Z[m['tag'] + 'events']
            assert np.array_equal(r.segs, np.concatenate([ev['start'], [ev['start'][-1] + ev['length'][-1]]]))
            assert r.corr_group == 'RawGenomeCorrected_000/BaseCalled_template' and r.strand == key[1]


@pytest.mark.gpu
def test_device_aggregation_matches_reference():
    from nanoraw_b200 import nanoraw_helper as nh
    rrc = nh.parse_fast5s(_write_corrected_files(), 'RawGenomeCorrected_000', ['BaseCalled_template'])
    cov = nh.get_coverage(rrc)
    L = META['genome_len']
    sizes = {'chrA': L, 'chrB': L, 'chrEmpty': 50}
    c2, means, sds, lens = nh.get_base_stats(rrc, sizes)
    assert np.isnan(means[('chrEmpty', '+')]).all() and (c2[('chrEmpty', '-')] == 0).all()
    for k, key in enumerate(TRACKS):
        assert np.array_equal(cov[key], Z['t%d/coverage' % k])
        assert np.array_equal(c2[key][:len(cov[key])], cov[key]) and (c2[key][len(cov[key]):] == 0).all()
        for got, name in ((means, 'means'), (sds, 'sds'), (lens, 'lengths')):
            gk = 't%d/%s' % (k, name)
            if gk not in Z.files:
                continue
            want = Z[gk]
            assert np.array_equal(np.isnan(got[key]), np.isnan(want))
            ok = ~np.isnan(want)
            # the reference adds in file order, the kernel in mapped-start order: float64 rounding only
            assert np.all(np.abs(got[key][ok] - want[ok]) <= 1e-12 * np.maximum(1.0, np.abs(want[ok])))
    assert np.array_equal(nh.get_base_means(rrc, sizes)[('chrA', '+')], means[('chrA', '+')], equal_nan=True)


@pytest.mark.gpu
def test_device_reads_events_match_reference():
    """get_reads_events (nanoraw_helper.py:446-484) against tests/golden/reads_events_v1.npz, made by the real
    reference over the same corrected reads: same positions, same event means per position (as multisets: the
    reference's order inside a position comes from an unstable argsort)"""
    from nanoraw_b200 import nanoraw_helper as nh
    G = np.load(os.path.join(HERE, 'golden', 'reads_events_v1.npz'))
    rrc = nh.parse_fast5s(_write_corrected_files(), 'RawGenomeCorrected_000', ['BaseCalled_template'])
    for k, key in enumerate(TRACKS):
        got = nh.get_reads_events(rrc[key], key[1] == '-')
        pos, off, vals = G['t%d/pos' % k], G['t%d/off' % k], G['t%d/values' % k]
        assert sorted(got) == pos.tolist()
        for i, p in enumerate(pos.tolist()):
            assert np.array_equal(np.sort(got[p]), vals[off[i]:off[i + 1]]), (key, p)
    assert nh.get_reads_events([], False) is None
