"""Size-independent properties of the CUDA path on batches far larger than the oracle can check
read by read (SURVEY.md section 4 invariants + invariances the arithmetic guarantees)."""
import numpy as np
import pytest

from nanoraw_b200 import synth
from nanoraw_b200.batch import ReadBatch

pytestmark = pytest.mark.gpu

N_READS = 600


@pytest.fixture(scope='module')
def engine():
    from nanoraw_b200.engine import ResquiggleEngine
    return ResquiggleEngine()


@pytest.fixture(scope='module')
def reads():
    cfgs = [synth.SynthConfig(span_mean=3000, span_sd=800, seed_base=61000),
            synth.SynthConfig(span_mean=2500, span_sd=500, del_rate=0.12, ins_rate=0.12,
                              clip_bases=(20, 200), seed_base=62000),
            synth.SynthConfig(span_mean=2000, span_sd=400, dwell_mean=6.4, seed_base=63000)]
    genome = synth.make_genome(1_000_000, seed=17)
    out = []
    for i in range(N_READS):
        cfg = cfgs[i % len(cfgs)]
        out.append(synth.make_read(genome, synth.kmer_levels(cfg), cfg, i))
    return out


@pytest.fixture(scope='module')
def result(engine, reads):
    from nanoraw_b200.engine import ResquiggleParams
    return engine.run(ReadBatch.from_reads(reads), ResquiggleParams())


def test_every_read_succeeds_and_segments_are_valid(reads, result):
    assert (result.status == 0).all()
    for r, read in enumerate(reads):
        segs = result.new_segs_of(r)
        g = np.frombuffer(read.genome_align, dtype=np.uint8)
        n_genome = int((g != ord('-')).sum())
        assert len(segs) == n_genome + 1 == int(result.n_bases[r]) + 1     # resquiggle.py:384-387
        assert segs[0] == 0 and segs[-1] == read.n_samples                  # :376-381
        assert np.diff(segs).min() >= 1                                     # :373-375
        ev = result.events_of(r)
        assert np.array_equal(ev['start'], segs[:-1]) and np.array_equal(ev['length'], np.diff(segs))
        assert ev['base'].tobytes() == g[g != ord('-')].tobytes()
        assert int(ev['length'].sum()) == read.n_samples
        # boundaries that are not basecaller boundaries were placed by the change-point search, which
        # keeps min_seg_len = 4 samples between neighbours (resquiggle.py:241-255)
        new = ~np.isin(segs, read.starts)
        idx = np.flatnonzero(new)
        assert (segs[idx] - segs[idx - 1] >= 4).all() and (segs[idx + 1] - segs[idx] >= 4).all()


def test_event_means_reassemble_the_signal_mean(engine, reads, result):
    """sum(mean_k * length_k) over a read == sum of its normalised signal"""
    from nanoraw_b200.engine import ResquiggleParams
    sub = list(range(0, N_READS, 25))
    db = engine.upload(ReadBatch.from_reads([reads[r] for r in sub]))
    _, sv, sigs = engine.normalize_signal(db, ResquiggleParams())
    for k, r in enumerate(sub):
        ev = result.events_of(r)
        total = float((ev['norm_mean'] * ev['length']).sum())
        assert abs(total - float(sigs[k].sum())) <= 1e-9 * max(1.0, np.abs(sigs[k]).sum())
        assert tuple(sv[k]) == tuple(result.scale_values[r])
        assert sigs[k].min() >= sv[k][2] and sigs[k].max() <= sv[k][3]     # winsorised to the limits


def test_integer_statistics_agree_with_numpy_order_kernel(engine, reads, result):
    from nanoraw_b200.engine import ResquiggleParams
    exact = engine.run(ReadBatch.from_reads(reads), ResquiggleParams(exact_stats=True))
    for r in range(N_READS):
        assert np.array_equal(exact.new_segs_of(r), result.new_segs_of(r))
        a, b = exact.events_of(r), result.events_of(r)
        assert np.max(np.abs(a['norm_mean'] - b['norm_mean'])) < 1e-12
        assert np.max(np.abs(a['norm_stdev'] - b['norm_stdev'])) < 1e-11


def test_results_do_not_depend_on_batch_composition(engine, reads, result):
    """reads are independent: reversing the batch order, or running a slice, changes nothing per read"""
    from nanoraw_b200.engine import ResquiggleParams
    rev = engine.run(ReadBatch.from_reads(reads[::-1]), ResquiggleParams())
    for r in range(0, N_READS, 7):
        q = N_READS - 1 - r
        assert np.array_equal(rev.new_segs_of(q), result.new_segs_of(r))
        assert np.array_equal(rev.events_of(q)['norm_mean'], result.events_of(r)['norm_mean'])
        assert np.array_equal(rev.scale_values[q], result.scale_values[r], equal_nan=True)
    again = engine.run(ReadBatch.from_reads(reads), ResquiggleParams())
    assert np.array_equal(again.status, result.status)
    for r in range(N_READS):  # bit-identical from run to run (no atomics on the data path)
        assert np.array_equal(again.new_segs_of(r), result.new_segs_of(r))
        assert np.array_equal(again.events_of(r), result.events_of(r))


def test_translation_and_power_of_two_scaling_invariance(engine, reads, result):
    """raw + c shifts `shift` by c and raw * 2 doubles shift and scale; the normalised signal is then
    bit-identical (all exact in float64), so boundaries and event statistics must not move"""
    from nanoraw_b200.engine import ResquiggleParams
    import copy
    sub = reads[:120]
    moved, doubled = [], []
    for rd in sub:
        a = copy.copy(rd)
        a.raw = (rd.raw.astype(np.int32) + 321).astype(np.int16)
        moved.append(a)
        b = copy.copy(rd)
        b.raw = (rd.raw.astype(np.int32) * 2).astype(np.int16)
        doubled.append(b)
    rm = engine.run(ReadBatch.from_reads(moved), ResquiggleParams())
    rd2 = engine.run(ReadBatch.from_reads(doubled), ResquiggleParams())
    for r in range(len(sub)):
        assert np.array_equal(rm.new_segs_of(r), result.new_segs_of(r))
        assert np.array_equal(rd2.new_segs_of(r), result.new_segs_of(r))
        sv, svm, svd = result.scale_values[r], rm.scale_values[r], rd2.scale_values[r]
        assert svm[0] == sv[0] + 321 and svm[1] == sv[1] and svd[0] == 2 * sv[0] and svd[1] == 2 * sv[1]
        assert np.array_equal(svm[2:], sv[2:]) and np.array_equal(svd[2:], sv[2:])
        ev = result.events_of(r)
        for other in (rm.events_of(r), rd2.events_of(r)):
            assert np.max(np.abs(other['norm_mean'] - ev['norm_mean'])) < 1e-12
            assert np.max(np.abs(other['norm_stdev'] - ev['norm_stdev'])) < 1e-11


def test_event_table_on_hand_made_segmentations(engine):
    """k_event_table keeps the prefixes of the last two 256-sample steps and resolves a round of boundaries only
    when 32 wait (or when some wait from the step before); it falls back to its step-by-step paths around clipped
    samples, bases of more than 768 samples and the ends of the read.  Segmentations made to sit on those seams --
    bursts of one- to three-sample bases (more than 32 boundaries per step for many steps), bases of hundreds to
    thousands of samples, spikes beside both -- must give the numpy-order kernel's table: start / length exactly,
    mean / sd within the contract (resquiggle.py:60-70)."""
    import ctypes as C
    import torch
    from nanoraw_b200 import _lib
    from nanoraw_b200.engine import DeviceResults, ResquiggleParams, _ptr
    cfg = synth.SynthConfig(span_mean=9000, span_sd=300, seed_base=64000)
    genome = synth.make_genome(1_000_000, seed=19)
    reads = [synth.make_read(genome, synth.kmer_levels(cfg), cfg, i) for i in range(12)]
    rng = np.random.default_rng(5)
    for rd in reads:  # spikes: isolated ones and short runs, some at the very ends of the read's samples
        raw = rd.raw.astype(np.int32)
        n = int(rd.starts[-1])
        r0 = int(rd.read_start_rel_to_raw)
        where = np.concatenate([rng.integers(0, n, 40), [0, 1, n - 1, n - 2], rng.integers(0, n - 8, 6)[:, None].repeat(5, 1).ravel()
                                + np.tile(np.arange(5), 6)])
        raw[r0 + where] += rng.choice([-3000, 3000], len(where))
        rd.raw = np.clip(raw, -32768, 32767).astype(np.int16)
    hb = ReadBatch.from_reads(reads)
    db = engine.upload(hb)
    pc = ResquiggleParams().to_c()
    out = DeviceResults(db.n_reads, db.total_align, engine.device)
    ws = engine.workspace(db.n_reads, db.total_align, pc)
    engine.run_device(db, pc, out=out, workspace=ws)
    torch.cuda.synchronize()
    assert bool((out.status[:db.n_reads] == 0).all())
    align_off = np.asarray(hb.align_off)
    segs = out.new_segs.cpu().numpy()
    n_bases = out.n_bases.cpu().numpy()

    def pattern(kind, n_samples, cap):
        """boundaries 0 = b[0] < b[1] < ... <= n_samples, at most cap bases"""
        b = [0]
        while b[-1] < n_samples and len(b) <= cap:
            left = n_samples - b[-1]
            budget = cap + 1 - len(b)
            if budget * 3000 < left:  # running out of bases: stride to the end
                b.append(b[-1] + min(left, 3000))
                continue
            mode = kind if kind != 'mix' else rng.choice(['burst', 'long', 'plain', 'edge'])
            if mode == 'burst':      # > 32 boundaries per 256-sample step, for a few steps
                d = rng.integers(1, 4, int(rng.integers(40, 400)))
            elif mode == 'long':     # around the 768-sample reach and the 1 024-sample limit of the 32-bit sums
                d = rng.choice([500, 700, 767, 768, 769, 1000, 1023, 1024, 1025, 1300, 2600], 1)
            elif mode == 'edge':     # boundaries on and beside the step edges
                base = (b[-1] // 256 + 1) * 256
                d = np.diff(np.concatenate([[b[-1]], base + np.array([-1, 0, 1, 255, 256, 257, 511, 513])]))
                d = d[d > 0]
            else:
                d = rng.geometric(1 / 7.9, int(rng.integers(20, 300))) + 1
            before = len(b)
            for step in d:
                if b[-1] + step > n_samples or len(b) > cap:
                    break
                b.append(int(b[-1] + step))
            if len(b) == before:     # nothing of this kind fits any more: one last base to the end
                b.append(n_samples)
        return np.array(b[:cap + 1], dtype=np.int32)

    kinds = ['mix', 'burst', 'long', 'edge', 'plain', 'mix']
    new_nb = n_bases.copy()
    for r in range(db.n_reads):
        a0, a1 = int(align_off[r]), int(align_off[r + 1])
        cap = a1 - a0 - 1
        bnd = pattern(kinds[r % len(kinds)], int(reads[r].starts[-1]), cap)
        nb = len(bnd) - 1
        assert 2 <= nb <= cap
        segs[a0 + r:a0 + r + nb + 1] = bnd
        new_nb[r] = nb
    out.new_segs.copy_(torch.from_numpy(segs))
    out.n_bases.copy_(torch.from_numpy(new_nb))

    tables = {}
    for flags in (0, 2):
        p = ResquiggleParams().to_c()
        p.debug_flags = flags
        o = DeviceResults(db.n_reads, db.total_align, engine.device)
        for name in ('norm_mean', 'norm_stdev'):
            getattr(o, name).fill_(-77.0)
        rc = engine.lib.nrq_event_table(C.byref(db.c), C.byref(p), _ptr(out.scale_values), _ptr(out.n_bases),
                                        _ptr(out.new_segs), _ptr(out.status), _ptr(o.norm_mean), _ptr(o.norm_stdev),
                                        _ptr(o.ev_start), _ptr(o.ev_length), _ptr(o.ev_base), engine._stream())
        _lib.check(rc, 'nrq_event_table')
        torch.cuda.synchronize()
        tables[flags] = {k: getattr(o, k).cpu().numpy() for k in ('norm_mean', 'norm_stdev', 'ev_start', 'ev_length')}
    fast, exact = tables[0], tables[2]
    for r in range(db.n_reads):
        a0 = int(align_off[r])
        nb = int(new_nb[r])
        sl = slice(a0, a0 + nb)
        bnd = segs[a0 + r:a0 + r + nb + 1]
        assert np.array_equal(fast['ev_start'][sl], bnd[:-1]) and np.array_equal(exact['ev_start'][sl], bnd[:-1])
        assert np.array_equal(fast['ev_length'][sl], np.diff(bnd)) and np.array_equal(exact['ev_length'][sl], np.diff(bnd))
        for col, tol in (('norm_mean', 1e-12), ('norm_stdev', 1e-10)):
            a, b = fast[col][sl], exact[col][sl]
            assert not np.isnan(b).any()
            assert np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))) < tol, (r, kinds[r % len(kinds)], col)


def test_long_reads_streamed_in_parts_give_the_same_table(engine):
    """nrq_params.long_read_parts: a read of more than 262 144 samples is cut at base boundaries into parts, one
    warp each (BASELINE configs[3]); a part is a read of its own from its first boundary on.  Same boundaries,
    same start / length, same means bit for bit (every path uses one formula for a base without clipped samples),
    standard deviations within a few ulps, and the whole table against the numpy-order kernel."""
    from nanoraw_b200.engine import ResquiggleParams
    genome = synth.make_genome(1_000_000, seed=23)
    cfg = synth.SynthConfig(span_mean=36_000, span_sd=2_000, seed_base=65000)
    reads = [synth.make_read(genome, synth.kmer_levels(cfg), cfg, i) for i in range(3)]
    reads += [synth.make_read(genome, synth.kmer_levels(synth.SynthConfig(span_mean=3000, seed_base=65100)),
                              synth.SynthConfig(span_mean=3000, seed_base=65100), 7)]  # a short one beside them
    assert max(int(r.starts[-1]) for r in reads) > 262_144
    batch = ReadBatch.from_reads(reads)
    whole = engine.run(batch, ResquiggleParams(long_read_parts=1))
    exact = engine.run(batch, ResquiggleParams(long_read_parts=1, exact_stats=True))
    for parts in (2, 5, 16, 64):
        cut = engine.run(batch, ResquiggleParams(long_read_parts=parts))
        for r in range(len(reads)):
            assert cut.ok(r) and whole.ok(r), (cut.error(r), whole.error(r))
            assert np.array_equal(cut.new_segs_of(r), whole.new_segs_of(r))
            a, b, c = cut.events_of(r), whole.events_of(r), exact.events_of(r)
            for col in ('start', 'length', 'base'):
                assert np.array_equal(a[col], b[col])
            assert np.array_equal(a['norm_mean'], b['norm_mean'])
            assert np.max(np.abs(a['norm_stdev'] - b['norm_stdev'])) < 1e-13
            assert np.max(np.abs(a['norm_mean'] - c['norm_mean'])) < 1e-12
            assert np.max(np.abs(a['norm_stdev'] - c['norm_stdev'])) < 1e-11
