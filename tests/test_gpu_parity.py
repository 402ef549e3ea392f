"""Parity of the CUDA path (through the C ABI) against the reference-pinned goldens and the
oracle.  Bit-exact: boundaries, scale values, normalised signal; event means / sds are also
compared exactly (the kernel follows numpy's summation order), with the contract's 1e-9
tolerance as the documented fallback bar."""
import os

import numpy as np
import pytest

from golden_util import load_golden
from nanoraw_b200 import synth
from nanoraw_b200.batch import ReadBatch
from oracle import resquiggle_oracle as orc

pytestmark = pytest.mark.gpu

INFO, CASES = load_golden()
TOL = 1e-9


def close(a, b):
    """abs(a-b) <= 1e-9 * max(1, |a|, |b|), NaN == NaN."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    lim = TOL * np.maximum(1.0, np.maximum(np.abs(a), np.abs(b)))
    return bool(np.all(both_nan | (np.abs(a - b) <= lim)))


@pytest.fixture(scope='module')
def engine():
    from nanoraw_b200.engine import ResquiggleEngine
    return ResquiggleEngine()


def given_scale_values(case):
    nt = case.params.get('norm_type', 'median')
    ch = INFO['channel']
    if nt == 'none':
        return [0.0, 1.0, np.nan, np.nan]
    if nt == 'pA_raw':
        return [-1.0 * np.float64(ch['offset']),
                np.float64(ch['digitisation']) / np.float64(ch['range']), np.nan, np.nan]
    return None


def groups_by_params():
    groups = {}
    for c in CASES:
        key = (c.params.get('norm_type', 'median'), c.params.get('outlier_thresh', 5),
               c.params.get('num_cpts_limit'), c.params.get('compute_sd', True))
        groups.setdefault(key, []).append(c)
    return groups


def run_cases(engine, cases, key, debug_flags=0, exact_stats=False):
    from nanoraw_b200.engine import ResquiggleParams
    norm_type, thresh, limit, sd = key
    sv_in = None
    if norm_type != 'median':
        sv_in = np.array([given_scale_values(c) for c in cases])
    batch = ReadBatch.from_reads([c.read for c in cases], sv_in)
    params = ResquiggleParams(norm_type=norm_type, outlier_thresh=thresh, num_cpts_limit=limit,
                              compute_sd=sd, debug_flags=debug_flags, exact_stats=exact_stats)
    return engine.run(batch, params)


def golden_new_segs(a):
    return np.concatenate([a['ev_start'], [a['ev_start'][-1] + a['ev_length'][-1]]]).astype(np.int64)


def check_case(case, res, r, exact_stats=False):
    from nanoraw_b200.errors import error_string
    a = case.a
    if case.error:
        assert error_string(res.status[r]) == case.error, case.name
        return
    assert res.ok(r), (case.name, res.error(r))
    assert np.array_equal(res.scale_values[r], a['scale_values'], equal_nan=True), case.name
    if 'ev_start' in a:
        assert np.array_equal(res.new_segs_of(r), golden_new_segs(a)), case.name
        ev = res.events_of(r)
        assert np.array_equal(ev['start'], a['ev_start']) and np.array_equal(ev['length'], a['ev_length'])
        assert np.array_equal(ev['base'], a['ev_base'])
        assert close(ev['norm_mean'], a['ev_mean']) and close(ev['norm_stdev'], a['ev_sd']), case.name
        if exact_stats:
            assert np.array_equal(ev['norm_mean'], a['ev_mean']), case.name + ' (mean not bit-exact)'
            assert np.array_equal(ev['norm_stdev'], a['ev_sd'], equal_nan=True), case.name + ' (sd not bit-exact)'
    else:
        # FAST5 write fails in the reference (outlier_threshold=None attr); compare change-points
        segs = res.new_segs_of(r)
        assert set(a['cpts'].tolist()) <= set(segs.tolist()), case.name
        assert len(segs) == int(res.n_bases[r]) + 1


@pytest.mark.parametrize('exact_stats', [False, True])
def test_golden_whole_path(engine, exact_stats):
    """default: integer-prefix event statistics (1e-9 contract); exact_stats: numpy summation order,
    compared bit for bit"""
    for key, cases in groups_by_params().items():
        res = run_cases(engine, cases, key, exact_stats=exact_stats)
        for r, case in enumerate(cases):
            check_case(case, res, r, exact_stats)


def test_golden_general_order_statistics(engine):
    """debug bit 0 forces the bisection path of nrq_norm_stats for every read."""
    for key, cases in groups_by_params().items():
        res = run_cases(engine, cases, key, debug_flags=1)
        for r, case in enumerate(cases):
            check_case(case, res, r)


def test_golden_indels(engine):
    cases = [c for c in CASES if 'indels' in c.a]
    db = engine.upload(ReadBatch.from_reads([c.read for c in cases]))
    status, n_ind, n_bases, off, indels, bases = engine.find_indels(db)
    for r, c in enumerate(cases):
        assert status[r] == 0
        mine = indels[off[r]:off[r + 1]]
        assert np.array_equal(mine[:, :3], c.a['indels']), c.name
        assert np.array_equal(mine[:, 3], np.concatenate([[0], np.cumsum(c.a['indels'][:, 2])[:-1]])
                              if len(mine) else mine[:, 3])
        g = np.frombuffer(c.read.genome_align, dtype=np.uint8)
        assert n_bases[r] == int((g != ord('-')).sum())
        a0 = int(db.host.align_off[r])
        assert np.array_equal(bases[a0:a0 + n_bases[r]], g[g != ord('-')])


def test_indel_scan_byte_path_equals_wide_path(engine):
    """nrq_find_indels takes 128-bit row loads when both alignment buffers sit on 16-byte boundaries and falls
    back to byte loads otherwise: same indels, bases and statuses either way (golden + seeded + malformed reads)"""
    import copy
    import torch
    from nanoraw_b200.engine import DeviceBatch
    reads = [c.read for c in CASES if 'indels' in c.a]
    reads += synth.make_reads(24, synth.SynthConfig(span_mean=700, span_sd=300, seed_base=61000),
                              genome=synth.make_genome(80_000, seed=3))
    reads += synth.make_reads(6, synth.config_indel_dense(seed_base=62000), genome=synth.make_genome(80_000, seed=4))
    bad = copy.copy(reads[3])
    bad.read_align = bad.read_align[:-1] + b'-'          # ends on a gap column
    reads.append(bad)
    db = engine.upload(ReadBatch.from_reads(reads))
    assert db.t['read_align'].data_ptr() % 16 == 0 and db.t['genome_align'].data_ptr() % 16 == 0
    wide = engine.find_indels(db)
    for shift in (1, 7, 8):
        t = dict(db.t)
        for name in ('read_align', 'genome_align'):
            buf = torch.zeros(db.t[name].numel() + 32, dtype=torch.uint8, device=db.t[name].device)
            buf[shift:shift + db.t[name].numel()] = db.t[name]
            t[name] = buf[shift:shift + db.t[name].numel()]
            assert t[name].data_ptr() % 16 == shift
        narrow = engine.find_indels(DeviceBatch.from_tensors(t, db.n_reads, db.total_align, db.device, host=db.host))
        for a, b in zip(wide, narrow):
            assert np.array_equal(a, b)
    status, n_ind, n_bases, off, indels, bases = wide
    assert status[-1] != 0 and (status[:-1] == 0).all()
    for r, rd in enumerate(reads[:-1]):
        want = np.array([tuple(i) for i in orc.get_all_indels(rd.read_align.decode(), rd.genome_align.decode())],
                        dtype=np.int64).reshape(-1, 3)
        assert np.array_equal(indels[off[r]:off[r + 1], :3], want), r
        g = np.frombuffer(rd.genome_align, dtype=np.uint8)
        a0 = int(db.host.align_off[r])
        assert np.array_equal(bases[a0:a0 + n_bases[r]], g[g != ord('-')]), r


def test_golden_normalised_signal(engine):
    from nanoraw_b200.engine import ResquiggleParams
    for key, cases in groups_by_params().items():
        cases = [c for c in cases if 'norm_signal' in c.a]
        if not cases:
            continue
        norm_type, thresh, _, _ = key
        sv_in = None if norm_type == 'median' else np.array([given_scale_values(c) for c in cases])
        db = engine.upload(ReadBatch.from_reads([c.read for c in cases], sv_in))
        st, sv, sigs = engine.normalize_signal(
            db, ResquiggleParams(norm_type=norm_type, outlier_thresh=thresh))
        for r, c in enumerate(cases):
            assert np.array_equal(sv[r], c.a['scale_values'], equal_nan=True), c.name
            assert np.array_equal(sigs[r], c.a['norm_signal']), c.name


SEEDED = {
    'standard_10kb': (synth.SynthConfig(seed_base=41000), 6),
    'short_4kb': (synth.config_short(42000), 8),
    'indel_dense': (synth.config_indel_dense(43000), 32),        # BASELINE configs[4]: 12 % / 12 % indels, 50-500 clips
    'low_dwell': (synth.SynthConfig(span_mean=3000, span_sd=300, dwell_mean=6.2, seed_base=44000), 6),
    'long_100kb': (synth.SynthConfig(span_mean=100_000, span_sd=5_000, seed_base=45000), 32),  # configs[3]
}


def _oracle_job(read):
    return orc.resquiggle_read(read.raw, read.read_start_rel_to_raw, read.starts, read.read_align.decode(),
                               read.genome_align.decode())


_ORACLE_CACHE = {}


def seeded_reads_and_oracle(name):
    """the seeded reads of one configuration and what the oracle makes of them (once per session; the oracle runs
    on all host cores: a 100 kb read takes it more than a second)"""
    if name not in _ORACLE_CACHE:
        import multiprocessing as mp
        cfg, n = SEEDED[name]
        reads = synth.make_reads(n, cfg, genome=synth.make_genome(1_000_000, seed=9))
        with mp.get_context('fork').Pool(min(len(reads), len(os.sched_getaffinity(0)))) as pool:
            outs = pool.map(_oracle_job, reads, chunksize=1)
        _ORACLE_CACHE[name] = (reads, outs)
    return _ORACLE_CACHE[name]


def assert_same_results(a, b, n_reads):
    assert np.array_equal(a.status, b.status) and np.array_equal(a.n_bases, b.n_bases)
    assert np.array_equal(a.n_indels, b.n_indels) and np.array_equal(a.region_samples, b.region_samples)
    assert np.array_equal(a.scale_values, b.scale_values, equal_nan=True)
    for r in range(n_reads):
        if a.ok(r):
            assert np.array_equal(a.new_segs_of(r), b.new_segs_of(r)), r
            assert np.array_equal(a.events_of(r), b.events_of(r)), r


@pytest.mark.parametrize('name', ['standard_10kb', 'indel_dense', 'low_dwell', 'tiny_batch'])
def test_speculative_search_changes_nothing(engine, name):
    """nrq_resegment runs the first change-point search of every indel group ahead of the read-by-read walk
    (k_speculate, one region per thread); debug bit 2 turns that off.  Everything must come out the same either
    way, and as in the oracle"""
    from nanoraw_b200.engine import ResquiggleParams
    genome = synth.make_genome(1_000_000, seed=9)
    if name == 'tiny_batch':
        reads = synth.make_reads(3, synth.SynthConfig(span_mean=150, span_sd=30, seed_base=71000), genome=genome)
    else:
        cfg, _ = SEEDED[name]
        reads = synth.make_reads(12, cfg, genome=genome)[:12]
    batch = ReadBatch.from_reads(reads)
    on = engine.run(batch, ResquiggleParams())
    off = engine.run(batch, ResquiggleParams(debug_flags=4))
    assert_same_results(on, off, len(reads))
    for r in (0, len(reads) - 1):
        out = orc.resquiggle_read(reads[r].raw, reads[r].read_start_rel_to_raw, reads[r].starts,
                                  reads[r].read_align.decode(), reads[r].genome_align.decode())
        assert on.ok(r) and np.array_equal(on.new_segs_of(r), out['new_segs'])


def test_speculative_search_at_the_edges_of_the_raw_array(engine):
    """k_speculate reads DAC codes as whole 16-byte words: regions that begin at the first sample of the raw
    array or end at its last one, and a raw array that starts off a 16-byte boundary, take the sample-by-sample
    path for the words that stick out"""
    import torch
    from nanoraw_b200.engine import DeviceBatch, ResquiggleParams
    rng = np.random.default_rng(7)
    reads = []
    for k in range(12):
        rd = handmade_read(['MI', 'MMD'] * 6, 100 + k, dwell=10, lead=1 + k % 3, tail=1 + k % 2)
        n = int(rd.starts[-1])
        rd.raw = rd.raw[25:25 + n].copy()      # no samples before or after the read
        rd.read_start_rel_to_raw = 0
        reads.append(rd)
    batch = ReadBatch.from_reads(reads)
    want = engine.run_via_torch(batch, ResquiggleParams(debug_flags=4))
    got = engine.run_via_torch(batch, ResquiggleParams())
    assert_same_results(got, want, len(reads))
    db = engine.upload(batch)
    for shift in (1, 3, 7):
        t = dict(db.t)
        buf = torch.zeros(db.t['raw'].numel() + 16, dtype=torch.int16, device=db.device)
        buf[shift:shift + db.t['raw'].numel()] = db.t['raw']
        t['raw'] = buf[shift:shift + db.t['raw'].numel()]
        assert t['raw'].data_ptr() % 16 == 2 * shift
        moved = DeviceBatch.from_tensors(t, db.n_reads, db.total_align, db.device, host=db.host)
        res = engine.download(moved, engine.run_device(moved, ResquiggleParams()))
        assert_same_results(res, want, len(reads))


@pytest.mark.parametrize('exact_stats', [False, True])
@pytest.mark.parametrize('name', list(SEEDED))
def test_seeded_reads_against_oracle(engine, name, exact_stats):
    from nanoraw_b200.engine import ResquiggleParams
    reads, outs = seeded_reads_and_oracle(name)
    res = engine.run(ReadBatch.from_reads(reads), ResquiggleParams(exact_stats=exact_stats))
    for r, read in enumerate(reads):
        out = outs[r]
        assert res.ok(r), res.error(r)
        assert res.scale_values_of(r) == tuple(float(v) for v in out['scale_values'])
        assert np.array_equal(res.new_segs_of(r), out['new_segs'])
        ev = res.events_of(r)
        for col in ('start', 'length', 'base'):
            assert np.array_equal(ev[col], out['events'][col])
        assert close(ev['norm_mean'], out['events']['norm_mean'])
        assert close(ev['norm_stdev'], out['events']['norm_stdev'])
        if exact_stats:
            assert np.array_equal(ev['norm_mean'], out['events']['norm_mean'])
            assert np.array_equal(ev['norm_stdev'], out['events']['norm_stdev'])
        else:  # integer-prefix statistics: a few ulps, far inside the contract
            assert np.max(np.abs(ev['norm_mean'] - out['events']['norm_mean'])) < 1e-13
            assert np.max(np.abs(ev['norm_stdev'] - out['events']['norm_stdev'])) < 1e-12


def handmade_read(units, seed, dwell=18, lead=40, tail=40):
    """a read whose alignment is `lead` matches, then `units` (strings over M / I / D, one alignment column
    each: match, base only in the read, base only in the genome), then `tail` matches"""
    rng = np.random.default_rng(seed)
    ops = 'M' * lead + ''.join(units) + 'M' * tail
    acgt = np.frombuffer(b'ACGT', dtype=np.uint8)
    read_row = np.where(np.array(list(ops)) == 'D', ord('-'), acgt[rng.integers(0, 4, len(ops))]).astype(np.uint8)
    genome_row = np.where(np.array(list(ops)) == 'I', ord('-'), acgt[rng.integers(0, 4, len(ops))]).astype(np.uint8)
    n_read = int((read_row != ord('-')).sum())
    dwells = rng.integers(dwell - 6, dwell + 7, n_read)
    starts = np.concatenate([[0], np.cumsum(dwells)]).astype(np.int64)
    levels = rng.normal(500, 60, n_read)
    raw = np.rint(np.repeat(levels, dwells) + rng.normal(0, 12, int(starts[-1]))).astype(np.int16)
    pre = rng.integers(300, 700, 25).astype(np.int16)
    return synth.SynthRead(raw=np.concatenate([pre, raw, pre]), read_start_rel_to_raw=25, starts=starts,
                           read_align=read_row.tobytes(), genome_align=genome_row.tobytes())


def test_groups_larger_than_the_planning_window(engine):
    """k_resegment plans indel groups 32 indels at a time: groups of more than 32 indels, groups that end exactly
    at a window edge and long chains of absorbed groups must come out as in the oracle"""
    from nanoraw_b200.engine import ResquiggleParams
    reads = [
        handmade_read(['MI', 'MD'] * 45, 1),                                  # one group of 90 indels
        handmade_read((['MI'] * 31 + ['M' * 12]) * 3 + ['MD'] * 33, 2),       # 31-, 31-, 31- and 33-indel groups
        handmade_read((['MMI'] * 8 + ['M' * 9]) * 12, 3),                     # 8-indel groups, 4 per window
        handmade_read(['MMMI', 'MMMD'] * 60, 4, dwell=9),                     # isolated indels that merge on retry
        handmade_read(['I' * 3 + 'M', 'D' * 2 + 'M'] * 40, 5, dwell=24),      # long runs, one open-ended group
        handmade_read(['MMMMMI', 'MMMMMMD'] * 40, 6, dwell=7),                # short dwell: retries and merges
    ]
    res = engine.run(ReadBatch.from_reads(reads), ResquiggleParams())
    for r, read in enumerate(reads):
        try:
            out = orc.resquiggle_read(read.raw, read.read_start_rel_to_raw, read.starts,
                                      read.read_align.decode(), read.genome_align.decode())
        except orc.RegionUnsatisfiable:
            from nanoraw_b200 import _lib
            assert res.status[r] == _lib.ST_UNSATISFIABLE, r
            continue
        assert res.ok(r), (r, res.error(r))
        assert np.array_equal(res.new_segs_of(r), out['new_segs']), r
        assert close(res.events_of(r)['norm_mean'], out['events']['norm_mean'])


def test_random_generator_settings_against_oracle(engine):
    """24 random settings of the read generator (indel rates from none to one base in three, dwell 6 to 24 samples,
    quantised to noisy signals, spike rates up to 2 %, narrow and wide DAC ranges), 6 reads each, whole path against
    the oracle: statuses, boundaries bit-exact, statistics inside the contract"""
    from nanoraw_b200 import _lib
    from nanoraw_b200.engine import ResquiggleParams
    rng = np.random.default_rng(20260101)
    genome = synth.make_genome(400_000, seed=5)
    reads = []
    for k in range(24):
        cfg = synth.SynthConfig(
            span_mean=float(rng.integers(300, 1800)), span_sd=50.0,
            del_rate=float(rng.choice([0.0, 0.02, 0.06, 0.12, 0.18])),
            ins_rate=float(rng.choice([0.0, 0.02, 0.06, 0.12, 0.18])),
            mismatch_rate=float(rng.uniform(0.0, 0.15)), indel_len_p=float(rng.uniform(0.4, 0.95)),
            dwell_mean=float(rng.uniform(6.0, 24.0)), dwell_min=int(rng.integers(1, 4)),
            level_sd=float(rng.choice([3.0, 25.0, 60.0, 400.0, 1500.0])),
            noise_sd=float(rng.choice([0.0, 0.7, 5.0, 12.0, 60.0])),
            spike_rate=float(rng.choice([0.0, 5e-4, 5e-3, 2e-2])), spike_sd=float(rng.choice([50.0, 400.0, 4000.0])),
            clip_bases=(0, int(rng.integers(0, 40))), seed_base=90000 + 100 * k)
        reads += synth.make_reads(6, cfg, genome=genome)
    res = engine.run(ReadBatch.from_reads(reads), ResquiggleParams())
    n_ok = 0
    for r, read in enumerate(reads):
        try:
            out = orc.resquiggle_read(read.raw, read.read_start_rel_to_raw, read.starts,
                                      read.read_align.decode(), read.genome_align.decode())
        except orc.RegionUnsatisfiable:
            assert res.status[r] == _lib.ST_UNSATISFIABLE, r
            continue
        except Exception as e:  # the reference's own failures, by message
            assert not res.ok(r) and res.error(r) == str(e), (r, res.error(r), str(e))
            continue
        assert res.ok(r), (r, res.error(r))
        n_ok += 1
        assert res.scale_values_of(r) == tuple(float(v) for v in out['scale_values']), r
        assert np.array_equal(res.new_segs_of(r), out['new_segs']), r
        ev = res.events_of(r)
        for col in ('start', 'length', 'base'):
            assert np.array_equal(ev[col], out['events'][col]), (r, col)
        assert close(ev['norm_mean'], out['events']['norm_mean']), r
        assert close(ev['norm_stdev'], out['events']['norm_stdev']), r
    assert n_ok >= len(reads) // 2


def test_unsatisfiable_region_terminates(engine):
    """mean dwell <= 5 samples/base: the reference loops forever (resquiggle.py:210-215);
    the device path must stop with NRQ_ST_UNSATISFIABLE like the oracle's guard."""
    from nanoraw_b200.engine import ResquiggleParams
    from nanoraw_b200 import _lib
    cfg = synth.SynthConfig(span_mean=300, span_sd=10, dwell_mean=3.2, dwell_min=2, seed_base=46000)
    reads = synth.make_reads(4, cfg, genome=synth.make_genome(100_000, seed=9))
    res = engine.run(ReadBatch.from_reads(reads), ResquiggleParams())
    for r, read in enumerate(reads):
        try:
            out = orc.resquiggle_read(read.raw, read.read_start_rel_to_raw, read.starts,
                                      read.read_align.decode(), read.genome_align.decode())
        except orc.RegionUnsatisfiable:
            assert res.status[r] == _lib.ST_UNSATISFIABLE
        else:
            assert res.ok(r) and np.array_equal(res.new_segs_of(r), out['new_segs'])


@pytest.mark.parametrize('chunk_bytes', [None, 150_000])
def test_host_buffer_entry_point(engine, chunk_bytes, monkeypatch):
    """nrq_resquiggle_batch_host (host pointers in/out) == the torch-managed route; with a small
    chunk size the batch flows through the multi-stream pipeline in many chunks."""
    import ctypes as C
    from nanoraw_b200 import _lib
    from nanoraw_b200.engine import ResquiggleParams
    if chunk_bytes:
        monkeypatch.setenv('NRQ_HOST_CHUNK_BYTES', str(chunk_bytes))
    reads = synth.make_reads(11, synth.config_short(47000), genome=synth.make_genome(200_000, seed=9))
    batch = ReadBatch.from_reads(reads)
    params = ResquiggleParams()
    want = engine.run(batch, params)
    lib = _lib.load()
    R, A = batch.n_reads, batch.total_align
    hb = _lib.NrqBatch()
    hb.n_reads = R
    keep = []
    for f in ReadBatch.FIELDS:
        arr = np.ascontiguousarray(getattr(batch, f))
        keep.append(arr)
        setattr(hb, f, arr.ctypes.data)
    outs = {'status': np.zeros(R, np.int32), 'n_bases': np.zeros(R, np.int32),
            'n_indels': np.zeros(R, np.int32), 'scale_values': np.zeros(4 * R),
            'new_segs': np.zeros(A + R, np.int32), 'norm_mean': np.zeros(A), 'norm_stdev': np.zeros(A),
            'ev_start': np.zeros(A, np.uint32), 'ev_length': np.zeros(A, np.uint32),
            'ev_base': np.zeros(A, np.uint8), 'region_samples': np.zeros(R, np.int32)}
    hr = _lib.NrqResults()
    for k, v in outs.items():
        setattr(hr, k, v.ctypes.data)
    h = lib.nrq_create(engine.device.index or 0)
    assert h
    pc = params.to_c()
    _lib.check(lib.nrq_resquiggle_batch_host(h, C.byref(hb), C.byref(pc), C.byref(hr)), 'host batch')
    lib.nrq_destroy(h)
    assert np.array_equal(outs['status'], want.status)
    assert np.array_equal(outs['n_bases'], want.n_bases)
    assert np.array_equal(outs['scale_values'].reshape(-1, 4), want.scale_values)
    for r in range(R):
        o, n = int(batch.align_off[r]), int(want.n_bases[r])
        assert np.array_equal(outs['new_segs'][o + r:o + r + n + 1], want.new_segs_of(r))
        assert np.array_equal(outs['norm_mean'][o:o + n], want.norm_mean[o:o + n])
        assert np.array_equal(outs['norm_stdev'][o:o + n], want.norm_stdev[o:o + n])
        assert np.array_equal(outs['ev_base'][o:o + n], want.ev_base[o:o + n])
        assert np.array_equal(outs['ev_start'][o:o + n], want.ev_start[o:o + n])



def test_host_entry_point_submit_and_wait(engine, monkeypatch):
    """three batches queued back to back through nrq_resquiggle_batch_host_submit, waited for in order and out of
    order, give what the synchronous call gives for each"""
    import ctypes as C
    from nanoraw_b200 import _lib
    from nanoraw_b200.engine import ResquiggleParams
    monkeypatch.setenv('NRQ_HOST_CHUNK_BYTES', '200000')  # several chunks per batch, slots shared across batches
    lib = _lib.load()
    params = ResquiggleParams()
    pc = params.to_c()
    genome = synth.make_genome(200_000, seed=9)
    h = lib.nrq_create(engine.device.index or 0)
    assert h
    jobs = []
    for k, n in enumerate((7, 12, 5)):
        reads = synth.make_reads(n, synth.config_short(49000 + 100 * k), genome=genome)
        batch = ReadBatch.from_reads(reads)
        R, A = batch.n_reads, batch.total_align
        hb = _lib.NrqBatch()
        hb.n_reads = R
        keep = [np.ascontiguousarray(getattr(batch, f)) for f in ReadBatch.FIELDS]
        for f, arr in zip(ReadBatch.FIELDS, keep):
            setattr(hb, f, arr.ctypes.data)
        outs = {'status': np.full(R, -1, np.int32), 'n_bases': np.zeros(R, np.int32), 'n_indels': np.zeros(R, np.int32),
                'scale_values': np.zeros(4 * R), 'new_segs': np.zeros(A + R, np.int32), 'norm_mean': np.zeros(A),
                'norm_stdev': np.zeros(A), 'ev_start': np.zeros(A, np.uint32), 'ev_length': np.zeros(A, np.uint32),
                'ev_base': np.zeros(A, np.uint8), 'region_samples': np.zeros(R, np.int32)}
        hr = _lib.NrqResults()
        for name, v in outs.items():
            setattr(hr, name, v.ctypes.data)
        ticket = lib.nrq_resquiggle_batch_host_submit(h, C.byref(hb), C.byref(pc), C.byref(hr))
        assert ticket == k
        jobs.append((batch, keep, hb, hr, outs, ticket))
    assert lib.nrq_resquiggle_batch_host_wait(h, 99) != 0          # unknown ticket
    for idx in (1, 0, 2):                                           # batches complete in submission order
        _lib.check(lib.nrq_resquiggle_batch_host_wait(h, jobs[idx][5]), 'wait')
    _lib.check(lib.nrq_resquiggle_batch_host_wait(h, 0), 'wait twice')
    lib.nrq_destroy(h)
    for batch, _, _, _, outs, _ in jobs:
        want = engine.run(batch, params)
        assert np.array_equal(outs['status'], want.status) and np.array_equal(outs['n_bases'], want.n_bases)
        for r in range(batch.n_reads):
            o, n = int(batch.align_off[r]), int(want.n_bases[r])
            assert np.array_equal(outs['new_segs'][o + r:o + r + n + 1], want.new_segs_of(r))
            assert np.array_equal(outs['norm_mean'][o:o + n], want.norm_mean[o:o + n])
            assert np.array_equal(outs['ev_length'][o:o + n], want.ev_length[o:o + n])


def test_host_path_equals_torch_path_and_pinned_batcher(engine):
    """engine.run (C-ABI host entry point) == the torch-tensor route == a PinnedBatcher fed read by read"""
    from nanoraw_b200.batch import PinnedBatcher
    from nanoraw_b200.engine import ResquiggleParams
    reads = synth.make_reads(9, synth.config_short(48000), genome=synth.make_genome(200_000, seed=9))
    batch = ReadBatch.from_reads(reads)
    params = ResquiggleParams()
    a = engine.run(batch, params)
    b = engine.run_via_torch(batch, params)
    pb = PinnedBatcher(max_reads=4, raw_capacity=10_000)
    for rd in reads:
        pb.add(rd.raw, rd.read_start_rel_to_raw, rd.starts, rd.read_align, rd.genome_align)
    c = engine.run_host(pb, params)
    for other in (b, c):
        assert np.array_equal(a.status, other.status) and np.array_equal(a.n_bases, other.n_bases)
        assert np.array_equal(a.scale_values, other.scale_values, equal_nan=True)
        for r in range(len(reads)):
            assert np.array_equal(a.new_segs_of(r), other.new_segs_of(r))
            assert np.array_equal(a.events_of(r), other.events_of(r))


def test_malformed_reads_get_a_status_and_do_not_disturb_their_neighbours(engine):
    """ragged / inconsistent inputs: every bad read is reported through status[r]; good reads in the same
    batch are unaffected"""
    import copy
    from nanoraw_b200 import _lib
    from nanoraw_b200.engine import ResquiggleParams
    cfg = synth.SynthConfig(span_mean=400, span_sd=40, seed_base=52000)
    good = synth.make_reads(6, cfg, genome=synth.make_genome(60_000, seed=9))
    want = engine.run(ReadBatch.from_reads(good), ResquiggleParams())
    bad = [copy.copy(r) for r in good]
    # 1: alignment starts with a gap column (parse_sam_record never produces this, resquiggle.py:623-637)
    bad[1].read_align = b'-' + bad[1].read_align[1:]
    # 2: one boundary too many for the read bases of the alignment
    bad[2].starts = np.concatenate([bad[2].starts, [bad[2].starts[-1] + 5]])
    # 3: read_start beyond the stored signal
    bad[3].read_start_rel_to_raw = len(bad[3].raw) + 10
    # 4: a zero-length basecaller event in the untouched part (resquiggle.py:373-375)
    s = bad[4].starts.copy()
    s[1] = s[0]
    bad[4].starts = s
    res = engine.run(ReadBatch.from_reads(bad), ResquiggleParams())
    assert res.status[1] == _lib.ST_BAD_ALIGNMENT
    assert res.status[2] == _lib.ST_BAD_INPUT
    assert res.status[3] in (_lib.ST_BAD_INPUT, _lib.ST_PAST_END)
    try:  # whatever the oracle makes of it (zero-length unless an indel group happens to replace that boundary)
        orc.resquiggle_read(bad[4].raw, bad[4].read_start_rel_to_raw, bad[4].starts, bad[4].read_align.decode(),
                            bad[4].genome_align.decode())
        assert res.status[4] == _lib.ST_OK
    except Exception as e:  # noqa: BLE001
        assert res.error(4) == str(e)
    for r in (0, 5):
        assert res.ok(r)
        assert np.array_equal(res.new_segs_of(r), want.new_segs_of(r))
        assert np.array_equal(res.events_of(r), want.events_of(r))
    # an empty batch is a no-op
    empty = engine.run(ReadBatch.from_arrays([], [], [], [], []), ResquiggleParams())
    assert empty.n_reads == 0
