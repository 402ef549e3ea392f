"""'pA' normalisation against vectors computed by the REAL reference (tests/golden/pa_v1.npz, made by
tests/golden/make_pa_golden.py): the oracle restatement on the CPU, the device path on the GPU."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
BLOB = np.load(os.path.join(HERE, 'golden', 'pa_v1.npz'))
META = json.loads(bytes(BLOB['meta']).decode())


def _case(c):
    k = c['key']
    kmers = [s.decode() for s in BLOB[k + '/kmers']]
    uniq = [s.decode() for s in BLOB[k + '/model_kmers']]
    model = {'mean': dict(zip(uniq, BLOB[k + '/model_mean'].tolist())),
             'inv_var': dict(zip(uniq, BLOB[k + '/model_inv_var'].tolist()))}
    return (BLOB[k + '/raw'], c['read_start_rel_to_raw'], c['n_samples'], c['outlier_thresh'], model,
            BLOB[k + '/event_means'], kmers, BLOB[k + '/norm_signal'], BLOB[k + '/scale_values'])


def _channel(mod):
    c = META['channel']
    return mod.channelInfo(c['offset'], c['range'], c['digitisation'], c['channel_number'], int(c['sampling_rate']))


@pytest.mark.parametrize('c', META['cases'], ids=[c['key'] for c in META['cases']])
def test_oracle_pa_matches_reference(c):
    from oracle import resquiggle_oracle as orc
    raw, start, n, thresh, model, ev_means, kmers, want, want_sv = _case(c)
    got, sv = orc.normalize_raw_signal(raw, start, n, 'pA', _channel(orc), thresh, pore_model=model,
                                       event_means=ev_means, event_kmers=kmers)
    assert np.array_equal(got, want)
    assert np.array_equal(np.array([np.nan if v is None else float(v) for v in sv]), want_sv, equal_nan=True)


@pytest.mark.gpu
@pytest.mark.parametrize('c', META['cases'], ids=[c['key'] for c in META['cases']])
def test_device_pa_matches_reference(c):
    from nanoraw_b200 import nanoraw_helper as nh
    raw, start, n, thresh, model, ev_means, kmers, want, want_sv = _case(c)
    got, sv = nh.normalize_raw_signal(raw, start, n, 'pA', _channel(nh), thresh, pore_model=model,
                                      event_means=ev_means, event_kmers=[k.encode() for k in kmers])
    assert np.array_equal(got, want)
    assert np.array_equal(np.array([np.nan if v is None else float(v) for v in sv]), want_sv, equal_nan=True)
