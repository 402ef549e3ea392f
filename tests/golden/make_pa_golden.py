#!/usr/bin/env python
"""'pA' normalisation (nanoraw_helper.py:184-203, 221-236, 248-262) computed by the REAL reference.

Runs the reference's own normalize_raw_signal (loaded from /root/reference by oracle/run_reference.py) with a
synthetic pore model and basecall event means, and freezes shift / scale / limits and the normalised signal into
tests/golden/pa_v1.npz, together with everything needed to call it again (the pore model as arrays).

    python tests/golden/make_pa_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from nanoraw_b200 import synth  # noqa: E402
from oracle import run_reference as rr  # noqa: E402

OUT = os.path.join(HERE, 'pa_v1.npz')
CHANNEL = {'offset': 12.0, 'range': 1467.61, 'digitisation': 8192.0, 'channel_number': '42', 'sampling_rate': 4000.0}


def main():
    _, ref_nh = rr.load_reference(stable_sort=True)
    genome = synth.make_genome(120_000, seed=21)
    chan = ref_nh.channelInfo(CHANNEL['offset'], CHANNEL['range'], CHANNEL['digitisation'],
                              CHANNEL['channel_number'], int(CHANNEL['sampling_rate']))
    blob, meta = {}, []
    for idx, (thresh, n_kmers) in enumerate(((5, 300), (None, 120), (2.5, 800))):
        rng = np.random.default_rng(300 + idx)
        cfg = synth.SynthConfig(span_mean=500, span_sd=30, seed_base=1300)
        read = synth.make_read(genome, synth.kmer_levels(cfg), cfg, idx)
        kmers = [bytes(k).decode() for k in rng.choice(list(b'ACGT'), size=(n_kmers, 5)).astype(np.uint8)]
        uniq = sorted(set(kmers))
        means = rng.normal(90, 12, len(uniq))
        inv_vars = 1.0 / rng.uniform(1.0, 4.0, len(uniq)) ** 2
        model = {'mean': dict(zip(uniq, means.tolist())), 'inv_var': dict(zip(uniq, inv_vars.tolist()))}
        ev_means = np.array([model['mean'][k] * 1.07 + 3.0 for k in kmers]) + rng.normal(0, 0.5, n_kmers)
        with np.errstate(all='raise'):
            norm, sv = ref_nh.normalize_raw_signal(
                read.raw, read.read_start_rel_to_raw, read.n_samples, 'pA', chan, thresh,
                pore_model=model, event_means=ev_means, event_kmers=kmers)
        key = 'p%d' % idx
        blob[key + '/raw'] = read.raw
        blob[key + '/kmers'] = np.array(kmers, dtype='S5')
        blob[key + '/model_kmers'] = np.array(uniq, dtype='S5')
        blob[key + '/model_mean'] = means
        blob[key + '/model_inv_var'] = inv_vars
        blob[key + '/event_means'] = ev_means
        blob[key + '/norm_signal'] = np.asarray(norm, dtype=np.float64)
        blob[key + '/scale_values'] = np.array([np.nan if v is None else float(v) for v in sv])
        meta.append({'key': key, 'read_start_rel_to_raw': int(read.read_start_rel_to_raw),
                     'n_samples': int(read.n_samples), 'outlier_thresh': thresh})
        print(key, thresh, tuple(sv))
    blob['meta'] = np.frombuffer(json.dumps({'numpy': np.__version__, 'channel': CHANNEL, 'cases': meta}).encode(),
                                 dtype=np.uint8)
    np.savez_compressed(OUT, **blob)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()
