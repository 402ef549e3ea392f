"""Generate tests/golden/hotpath_v1.npz by running the REAL reference.

Run in the build container (needs /root/reference):
    python tests/golden/make_golden.py

Every output stored here was computed by the reference's own functions
(nanoraw_helper.normalize_raw_signal, resquiggle.get_indel_groups,
resquiggle.resquiggle_read + write_new_fast5_group) loaded through
oracle/run_reference.py; the oracle and the CUDA path are then tested against
this file on machines where the reference does not exist (the GPU box).

Two result sets per read:
  *_stable  -- reference with running_diffs.argsort(kind='stable')  (contract)
  *_default -- reference exactly as written (numpy default argsort on the
               generating host; informative, tie order is host dependent)
"""
import copy
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from nanoraw_b200 import synth  # noqa: E402
from oracle import run_reference as rr  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden', 'hotpath_v1.npz')


def low_complexity_genome(length, seed):
    """homopolymer / short-repeat rich genome: exercises the ambiguity
    extension of resquiggle.py:187-195."""
    rng = np.random.default_rng(seed)
    out = []
    total = 0
    acgt = np.frombuffer(b'ACGT', dtype=np.uint8)
    while total < length:
        kind = rng.integers(0, 3)
        if kind == 0:   # homopolymer
            unit = acgt[rng.integers(0, 4, size=1)]
        elif kind == 1:  # di/tri-nucleotide repeat
            unit = acgt[rng.integers(0, 4, size=rng.integers(2, 4))]
        else:
            unit = acgt[rng.integers(0, 4, size=rng.integers(3, 9))]
        reps = int(rng.integers(1, 7))
        piece = np.tile(unit, reps)
        out.append(piece)
        total += len(piece)
    return np.concatenate(out)[:length]


def edge_indel_read(base):
    """clean read -> deletion at column 1, insertion before the last column."""
    r = copy.deepcopy(base)
    ra = bytearray(r.read_align)
    ga = bytearray(r.genome_align)
    assert ra == ga
    starts = list(r.starts)
    # genome base 1 missing from the read
    ra[1] = ord('-')
    del starts[1]
    # extra read base just before the last column
    ra.insert(len(ra) - 1, ord('T' if ga[-2] != ord('T') else 'G'))
    ga.insert(len(ga) - 1, ord('-'))
    lo, hi = starts[-2], starts[-1]
    assert hi - lo >= 2
    starts.insert(len(starts) - 1, lo + (hi - lo) // 2)
    r.read_align, r.genome_align = bytes(ra), bytes(ga)
    r.starts = np.asarray(starts, dtype=np.int64)
    return r


def build_cases():
    cases = []
    small_genome = synth.make_genome(300_000, seed=7)
    lc_genome = low_complexity_genome(300_000, seed=11)

    def add(name, cfg, index, genome=small_genome, mutate=None, **params):
        levels = synth.kmer_levels(cfg)
        read = synth.make_read(genome, levels, cfg, index)
        if mutate is not None:
            read = mutate(read)
        cases.append((name, read, params))

    C = synth.SynthConfig
    for i, span in enumerate((300, 500, 700, 900, 1100, 1300, 1500, 1501)):
        add('standard_%d' % span, C(span_mean=span, span_sd=1, seed_base=100), i)
    for i in range(4):
        add('indel_dense_%d' % i, C(span_mean=700, span_sd=60, del_rate=0.12,
                                    ins_rate=0.12, clip_bases=(5, 40), seed_base=200), i)
    for i in range(4):
        add('low_dwell_%d' % i, C(span_mean=600, span_sd=50, dwell_mean=6.5,
                                  seed_base=300), i)
    for i in range(4):
        add('homopolymer_%d' % i, C(span_mean=800, span_sd=50, seed_base=400), i,
            genome=lc_genome)
    for i in range(2):
        add('clean_%d' % i, C(span_mean=400, span_sd=20, del_rate=0, ins_rate=0,
                              mismatch_rate=0, seed_base=500), i)
    for i in range(2):
        add('spiky_%d' % i, C(span_mean=800, span_sd=50, spike_rate=0.02,
                              seed_base=600), i)
    for i in range(2):
        add('long_indels_%d' % i, C(span_mean=900, span_sd=50, indel_len_p=0.3,
                                    seed_base=700), i)
    add('norm_none', C(span_mean=600, span_sd=10, seed_base=800), 0, norm_type='none')
    add('norm_none_noclip', C(span_mean=600, span_sd=10, seed_base=800), 1,
        norm_type='none', outlier_thresh=None)
    add('norm_pA_raw', C(span_mean=600, span_sd=10, seed_base=800), 2,
        norm_type='pA_raw')
    add('norm_pA_raw_noclip', C(span_mean=600, span_sd=10, seed_base=800), 3,
        norm_type='pA_raw', outlier_thresh=None)
    add('median_noclip', C(span_mean=600, span_sd=10, seed_base=800), 4,
        outlier_thresh=None)
    add('median_thresh_2p5', C(span_mean=600, span_sd=10, seed_base=800), 5,
        outlier_thresh=2.5)
    add('skip_sd', C(span_mean=600, span_sd=10, seed_base=800), 6, compute_sd=False)
    add('cpts_limit_3', C(span_mean=600, span_sd=10, seed_base=800), 7,
        num_cpts_limit=3)
    add('cpts_limit_100', C(span_mean=600, span_sd=10, seed_base=800), 8,
        num_cpts_limit=100)

    def constant(r):
        r.raw = np.full_like(r.raw, 500)
        return r

    def half_constant(r):
        raw = r.raw.copy()
        raw[::2] = 500
        raw[1::4] = 500
        r.raw = raw
        return r

    def truncated(r):
        r.raw = r.raw[:r.read_start_rel_to_raw + r.n_samples - 37]
        return r

    def wide(r):
        rng = np.random.default_rng(99)
        r.raw = rng.integers(-32768, 32768, size=len(r.raw)).astype(np.int16)
        return r

    def wide_outliers(r):
        # normal signal plus a few extreme samples far outside any window
        raw = r.raw.copy()
        raw[r.read_start_rel_to_raw + 5] = 32767
        raw[r.read_start_rel_to_raw + 77] = -32768
        raw[r.read_start_rel_to_raw + 78] = 20000
        r.raw = raw
        return r

    def quantised(r):
        # heavily quantised signal: many exact ties in the running differences
        r.raw = ((r.raw // 16) * 16).astype(np.int16)
        return r

    add('fail_constant', C(span_mean=400, span_sd=10, seed_base=900), 0, mutate=constant)
    add('fail_half_constant', C(span_mean=400, span_sd=10, seed_base=900), 1,
        mutate=half_constant)
    add('fail_truncated', C(span_mean=400, span_sd=10, seed_base=900), 2, mutate=truncated)
    add('wide_uniform', C(span_mean=500, span_sd=10, seed_base=900), 3, mutate=wide)
    add('wide_outliers', C(span_mean=500, span_sd=10, seed_base=900), 4,
        mutate=wide_outliers)
    for i in range(3):
        add('quantised_%d' % i, C(span_mean=900, span_sd=50, seed_base=1000), i,
            mutate=quantised)
    add('edge_indels', C(span_mean=300, span_sd=5, del_rate=0, ins_rate=0,
                         mismatch_rate=0, seed_base=1100), 0, mutate=edge_indel_read)
    return cases


CHANNEL = {'offset': 12.0, 'range': 1467.61, 'digitisation': 8192.0,
           'channel_number': '42', 'sampling_rate': 4000.0}


def run_case(ref_rsq, ref_nh, tag, read, params):
    """full route through the reference's resquiggle_read against a mem FAST5;
    falls back to the function-level route when only the FAST5 write fails."""
    norm_type = params.get('norm_type', 'median')
    outlier_thresh = params.get('outlier_thresh', 5)
    limit = params.get('num_cpts_limit')
    compute_sd = params.get('compute_sd', True)
    chan = ref_nh.channelInfo(CHANNEL['offset'], CHANNEL['range'],
                              CHANNEL['digitisation'], CHANNEL['channel_number'],
                              int(CHANNEL['sampling_rate']))
    out = {'error': '', 'error_full': ''}
    # function-level route (always)
    try:
        norm, sv, groups = rr.run_hot_path(
            ref_rsq, ref_nh, read, norm_type, outlier_thresh, None, limit,
            compute_sd, chan)
        out['scale_values'] = np.array(
            [np.nan if v is None else float(v) for v in sv], dtype=np.float64)
        out['norm_signal'] = np.asarray(norm, dtype=np.float64)
        out['indels'] = np.array(
            [tuple(i) for g in groups for i in g.indels], dtype=np.int64).reshape(-1, 3)
        out['groups'] = np.array([(g.start, g.end, len(g.cpts)) for g in groups],
                                 dtype=np.int64).reshape(-1, 3)
        out['cpts'] = np.array([c for g in groups for c in g.cpts], dtype=np.int64)
    except Exception as e:  # noqa: BLE001
        out['error'] = '%s: %s' % (type(e).__name__, e)
    # full route (resquiggle_read + write_new_fast5_group)
    name = rr.make_mem_fast5('mem://golden/' + tag, read, CHANNEL)
    try:
        sub = rr.run_resquiggle_read(ref_rsq, name, read, norm_type, outlier_thresh,
                                     None, limit, compute_sd)
        ev = sub['Events'].value
        out['ev_mean'] = ev['norm_mean'].copy()
        out['ev_sd'] = ev['norm_stdev'].copy()
        out['ev_start'] = ev['start'].copy()
        out['ev_length'] = ev['length'].copy()
        out['ev_base'] = ev['base'].copy()
        out['read_segments'] = sub['Alignment']['read_segments'].value.copy()
        attrs = dict(sub.attrs)
        out['attr_scale_values'] = np.array(
            [attrs['shift'], attrs['scale'], attrs['lower_lim'], attrs['upper_lim']],
            dtype=np.float64)
    except Exception as e:  # noqa: BLE001
        out['error_full'] = '%s: %s' % (type(e).__name__, e)
    return out


def main():
    ref_d, nh_d = rr.load_reference(stable_sort=False)
    ref_s, nh_s = rr.load_reference(stable_sort=True)
    cases = build_cases()
    blob = {}
    meta = []
    for idx, (name, read, params) in enumerate(cases):
        res_s = run_case(ref_s, nh_s, name + '_s', read, params)
        res_d = run_case(ref_d, nh_d, name + '_d', read, params)
        key = 'c%02d' % idx
        blob[key + '/raw'] = read.raw
        blob[key + '/starts'] = np.asarray(read.starts, dtype=np.int64)
        blob[key + '/read_align'] = np.frombuffer(read.read_align, dtype=np.uint8)
        blob[key + '/genome_align'] = np.frombuffer(read.genome_align, dtype=np.uint8)
        for field, val in res_s.items():
            if field in ('error', 'error_full'):
                continue
            if field == 'norm_signal' and idx >= 4 and not name.startswith('norm_'):
                continue  # keep the fixture small: full signal for a few reads only
            blob[key + '/' + field] = val
        for field in ('cpts', 'ev_start', 'ev_length'):
            if field in res_d:
                blob[key + '/default_' + field] = res_d[field]
        meta.append({
            'key': key, 'name': name, 'params': params,
            'read_start_rel_to_raw': int(read.read_start_rel_to_raw),
            'error': res_s['error'], 'error_full': res_s['error_full'],
            'error_default': res_d['error'],
            'n_bases': int(len(read.starts) - 1), 'n_samples': read.n_samples,
            'default_differs': bool(
                'cpts' in res_s and 'cpts' in res_d and
                not np.array_equal(res_s['cpts'], res_d['cpts'])),
        })
        print(idx, name, params, 'err=%r full=%r' % (res_s['error'], res_s['error_full']),
              'groups', len(res_s.get('groups', [])), 'default_differs',
              meta[-1]['default_differs'])
    blob['meta'] = np.frombuffer(json.dumps(
        {'numpy': np.__version__, 'channel': CHANNEL, 'cases': meta}).encode(),
        dtype=np.uint8)
    np.savez_compressed(OUT, **blob)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()
