"""The reference's call surface (normalize_raw_signal, get_indel_groups, resquiggle_read,
write_new_fast5_group, resquiggle_worker) served by the CUDA path, checked against vectors the
REAL reference produced (tests/golden/hotpath_v1.npz)."""
import queue

import numpy as np
import pytest

from golden_util import load_golden
from nanoraw_b200 import fast5_io, synth

pytestmark = pytest.mark.gpu

INFO, CASES = load_golden()
OK_CASES = [c for c in CASES if not c.error]
CHANNEL = INFO['channel']


def make_fast5(name, read, with_corrected_group=True):
    f = fast5_io.MemFile(name, 'w')
    rd = f.create_group('Raw/Reads/Read_1')
    rd.create_dataset('Signal', data=np.asarray(read.raw, dtype=np.int16))
    rd.attrs['read_id'] = 'r'
    rd.attrs['start_time'] = np.uint64(0)
    ch = f.create_group('UniqueGlobalKey/channel_id')
    for k in ('offset', 'range', 'digitisation', 'sampling_rate'):
        ch.attrs[k] = np.float64(CHANNEL[k])
    ch.attrs['channel_number'] = CHANNEL['channel_number']
    f.create_group('Analyses/Basecall_1D_000')
    if with_corrected_group:
        g = f.create_group('Analyses/RawGenomeCorrected_000')
        g.attrs['nanoraw_version'] = '0.5'
        g.attrs['basecall_group'] = 'Basecall_1D_000'
    f.close()
    return name


def channel_info():
    from nanoraw_b200 import nanoraw_helper as nh
    return nh.channelInfo(np.float64(CHANNEL['offset']), np.float64(CHANNEL['range']),
                          np.float64(CHANNEL['digitisation']), CHANNEL['channel_number'],
                          int(CHANNEL['sampling_rate']))


def read_info_for(read):
    from nanoraw_b200 import resquiggle as rsq
    n_ins, n_del, n_match, n_mm = read.tallies()
    return (rsq.readInfo('r', 'BaseCalled_template', 0, 0, n_ins, n_del, n_match, n_mm),
            rsq.genomeLoc(1234, '+', 'chrA'))


@pytest.mark.parametrize('case', [c for c in CASES if 'norm_signal' in c.a], ids=lambda c: c.name)
def test_normalize_raw_signal(case):
    from nanoraw_b200 import nanoraw_helper as nh
    kw = case.kwargs()
    sig, sv = nh.normalize_raw_signal(
        case.read.raw, case.read.read_start_rel_to_raw, case.read.n_samples, kw['norm_type'],
        channel_info(), kw['outlier_thresh'])
    want = case.a['scale_values']
    got = np.array([np.nan if v is None else float(v) for v in sv])
    assert np.array_equal(got, want, equal_nan=True)
    assert np.array_equal(sig, case.a['norm_signal'])


def test_normalize_with_explicit_shift_scale_and_limits():
    """the plotting call (plot_commands.py:219-222): shift/scale/limits handed in"""
    from nanoraw_b200 import nanoraw_helper as nh
    from oracle import resquiggle_oracle as orc
    case = OK_CASES[0]
    r = case.read
    want, wsv = orc.normalize_raw_signal(r.raw, r.read_start_rel_to_raw, r.n_samples,
                                         shift=498.5, scale=41.25, lower_lim=-3.0, upper_lim=2.5)
    got, gsv = nh.normalize_raw_signal(r.raw, r.read_start_rel_to_raw, r.n_samples,
                                       shift=498.5, scale=41.25, lower_lim=-3.0, upper_lim=2.5)
    assert np.array_equal(got, want) and tuple(gsv) == tuple(wsv)


def test_normalize_constant_signal_raises_like_numpy():
    from nanoraw_b200 import nanoraw_helper as nh
    case = [c for c in CASES if c.name == 'fail_constant'][0]
    with pytest.raises(FloatingPointError) as ei:
        nh.normalize_raw_signal(case.read.raw, case.read.read_start_rel_to_raw,
                                case.read.n_samples, 'median', None, 5)
    assert 'FloatingPointError: %s' % ei.value == case.error


@pytest.mark.parametrize('case', [c for c in OK_CASES if 'norm_signal' in c.a], ids=lambda c: c.name)
def test_get_indel_groups(case):
    from nanoraw_b200 import resquiggle as rsq
    r = case.read
    groups = rsq.get_indel_groups(r.align_vals, r.starts, case.a['norm_signal'], 4, None,
                                  case.kwargs()['num_cpts_limit'])
    want = case.a['groups']
    assert [(g.start, g.end, len(g.cpts)) for g in groups] == [tuple(int(v) for v in w) for w in want]
    assert [c for g in groups for c in g.cpts] == case.a['cpts'].tolist()
    assert [tuple(i) for g in groups for i in g.indels] == [tuple(int(v) for v in w) for w in case.a['indels']]


def test_get_indel_groups_cpts_limit_raises():
    from nanoraw_b200 import resquiggle as rsq
    case = [c for c in OK_CASES if 'norm_signal' in c.a][0]
    with pytest.raises(RuntimeError) as ei:
        rsq.get_indel_groups(case.read.align_vals, case.read.starts, case.a['norm_signal'], 4, None, 1)
    assert str(ei.value) == 'Reached maximum number of changepoints for a single indel'


@pytest.mark.parametrize('case', CASES, ids=lambda c: c.name)
def test_resquiggle_read_writes_reference_layout(case):
    from nanoraw_b200 import resquiggle as rsq
    r = case.read
    name = make_fast5('mem://host_api/' + case.name, r)
    info, loc = read_info_for(r)
    kw = case.kwargs()
    args = (name, r.read_start_rel_to_raw, r.starts, kw['norm_type'], kw['outlier_thresh'],
            r.align_vals, None, kw['num_cpts_limit'], loc, info, 'Basecall_1D_000',
            'RawGenomeCorrected_000', kw['compute_sd'], None)
    full_err = case.meta['error_full']
    if full_err:
        with pytest.raises(Exception) as ei:
            rsq.resquiggle_read(*args)
        assert '%s: %s' % (type(ei.value).__name__, ei.value) == full_err
        return
    rsq.resquiggle_read(*args)
    sub = fast5_io.MemFile(name)['/Analyses/RawGenomeCorrected_000/BaseCalled_template']
    a = case.a
    got_sv = np.array([sub.attrs[k] for k in ('shift', 'scale', 'lower_lim', 'upper_lim')], dtype=np.float64)
    assert np.array_equal(got_sv, a['attr_scale_values'])
    assert sub.attrs['norm_type'] == kw['norm_type'] and sub.attrs['outlier_threshold'] == kw['outlier_thresh']
    ev = sub['Events'][()]
    assert ev.dtype == np.dtype(rsq.EVENT_DTYPE)
    assert np.array_equal(ev['start'], a['ev_start']) and np.array_equal(ev['length'], a['ev_length'])
    assert np.array_equal(ev['base'], a['ev_base'])
    lim = 1e-9 * np.maximum(1.0, np.abs(a['ev_mean']))
    assert np.all(np.abs(ev['norm_mean'] - a['ev_mean']) <= lim)
    if kw['compute_sd']:
        assert np.all(np.abs(ev['norm_stdev'] - a['ev_sd']) <= 1e-9 * np.maximum(1.0, a['ev_sd']))
    else:
        assert np.isnan(ev['norm_stdev']).all()
    assert sub['Events'].attrs['read_start_rel_to_raw'] == r.read_start_rel_to_raw
    al = sub['Alignment']
    assert al['read_alignment'][()].tobytes() == r.read_align
    assert al['genome_alignment'][()].tobytes() == r.genome_align
    assert np.array_equal(al['read_segments'][()], r.starts)
    assert al.attrs['mapped_start'] == 1234 and al.attrs['mapped_strand'] == '+'
    assert al.attrs['num_insertions'] == info.Insertions and al.attrs['num_matches'] == info.Matches


def test_write_new_fast5_group_computes_events_on_device():
    """the reference signature: norm_signal + new_segs in, event statistics computed inside"""
    from nanoraw_b200 import nanoraw_helper as nh
    from nanoraw_b200 import resquiggle as rsq
    case = [c for c in OK_CASES if 'norm_signal' in c.a and 'ev_mean' in c.a][0]
    r = case.read
    name = make_fast5('mem://host_api/write_only', r)
    info, loc = read_info_for(r)
    a = case.a
    new_segs = np.concatenate([a['ev_start'], [a['ev_start'][-1] + a['ev_length'][-1]]]).astype(np.int64)
    sv = nh.scaleValues(*a['scale_values'])
    rsq.write_new_fast5_group(
        name, loc, info, r.read_start_rel_to_raw, new_segs, a['ev_base'].tobytes().decode(),
        r.align_vals, r.starts, a['norm_signal'], sv, 'RawGenomeCorrected_000',
        'BaseCalled_template', 'median', 5, True)
    ev = fast5_io.MemFile(name)['/Analyses/RawGenomeCorrected_000/BaseCalled_template/Events'][()]
    assert np.array_equal(ev['norm_mean'], a['ev_mean'])      # numpy summation order: bit-exact
    assert np.array_equal(ev['norm_stdev'], a['ev_sd'])
    assert np.array_equal(ev['start'], a['ev_start'])


def test_resquiggle_worker_batches_reads_and_routes_failures():
    from nanoraw_b200 import resquiggle as rsq
    good = [c for c in OK_CASES if c.kwargs() == OK_CASES[0].kwargs()][:6]
    bad = [c for c in CASES if c.name in ('fail_constant', 'fail_truncated')]
    q_in, q_fail = queue.Queue(), queue.Queue()
    names = {}
    for c in good + bad:
        name = make_fast5('mem://worker/' + c.name, c.read)
        names[c.name] = name
        info, loc = read_info_for(c.read)
        q_in.put((name, [(c.read.align_vals, loc, c.read.starts, c.read.read_start_rel_to_raw, info)]))
    # one file with two subgroups (2D read): processed in separate batches, same file
    c2 = good[0]
    name2 = make_fast5('mem://worker/two_subgroups', c2.read)
    info, loc = read_info_for(c2.read)
    info_c = info._replace(Subgroup='BaseCalled_complement')
    q_in.put((name2, [(c2.read.align_vals, loc, c2.read.starts, c2.read.read_start_rel_to_raw, info),
                      (c2.read.align_vals, loc, c2.read.starts, c2.read.read_start_rel_to_raw, info_c)]))
    q_in.put((None, None))
    rsq.resquiggle_worker(q_in, q_fail, 'Basecall_1D_000', 'RawGenomeCorrected_000', 'median', 5,
                          None, None, True, None)
    failures = {}
    while not q_fail.empty():
        err, rid = q_fail.get()
        failures[rid] = err
    assert failures == {
        'BaseCalled_template :: ' + names['fail_constant']: 'invalid value encountered in divide',
        'BaseCalled_template :: ' + names['fail_truncated']: 'New segments end past raw signal values.'}
    for c in good:
        ev = fast5_io.MemFile(names[c.name])[
            '/Analyses/RawGenomeCorrected_000/BaseCalled_template/Events'][()]
        assert np.array_equal(ev['start'], c.a['ev_start'])
    grp = fast5_io.MemFile(name2)['/Analyses/RawGenomeCorrected_000']
    assert sorted(grp.keys()) == ['BaseCalled_complement', 'BaseCalled_template']
    assert np.array_equal(grp['BaseCalled_complement/Events'][()]['start'], c2.a['ev_start'])


def test_pa_normalisation_against_oracle():
    """'pA': k-mer model fit (nanoraw_helper.py:184-203, 228-236) feeding the device path"""
    from nanoraw_b200 import nanoraw_helper as nh
    from oracle import resquiggle_oracle as orc
    rng = np.random.default_rng(5)
    case = OK_CASES[1]
    r = case.read
    kmers = [bytes(k) for k in rng.choice(list(b'ACGT'), size=(200, 5)).astype(np.uint8)]
    model = {'mean': {}, 'inv_var': {}}
    for k in set(kmers):
        model['mean'][k.decode()] = float(rng.normal(90, 12))
        model['inv_var'][k.decode()] = float(1.0 / rng.uniform(1.0, 4.0) ** 2)
    ev_means = np.array([model['mean'][k.decode()] * 1.07 + 3.0 + rng.normal(0, 0.5) for k in kmers])
    chan = channel_info()
    ochan = orc.channelInfo(*chan)
    omodel = {'mean': model['mean'], 'inv_var': model['inv_var']}
    want, wsv = orc.normalize_raw_signal(r.raw, r.read_start_rel_to_raw, r.n_samples, 'pA', ochan, 5,
                                         pore_model=omodel, event_means=ev_means,
                                         event_kmers=[k.decode() for k in kmers])
    got, gsv = nh.normalize_raw_signal(r.raw, r.read_start_rel_to_raw, r.n_samples, 'pA', chan, 5,
                                       pore_model=model, event_means=ev_means, event_kmers=kmers)
    assert np.array_equal(got, want)
    assert tuple(float(v) for v in gsv) == tuple(float(v) for v in wsv)


def test_running_difference_trace():
    """f-4: plot_correction's trace (plot_commands.py:225-229) is bit-identical to numpy's"""
    from nanoraw_b200 import resquiggle as rsq
    from oracle import resquiggle_oracle as orc
    case = [c for c in OK_CASES if 'norm_signal' in c.a][0]
    for lo, width in ((0, 200), (137, 501), (50, 8), (10, 7)):
        window = case.a['norm_signal'][lo:lo + width]
        got = rsq.running_differences(window)
        want = orc.running_diffs(window) if width >= 8 else np.zeros(0)
        assert np.array_equal(got, want)


def test_native_front_end_feeds_the_device_batch():
    """f-2 through the C ABI end to end: SAM records and basecall events go through libnrqio
    (nrqio_events_to_starts, nrqio_frontend_batch), whose CSR arrays are handed to nrq_resquiggle_batch_host as they
    are; the result must equal what the per-read Python front end + engine.run give"""
    import ctypes as C
    from nanoraw_b200 import _lib, _libio, frontend, synth
    from nanoraw_b200.batch import ReadBatch
    from nanoraw_b200.engine import ResquiggleEngine, ResquiggleParams
    engine = ResquiggleEngine()
    genome = synth.make_genome(300_000, seed=12)
    contig = genome.tobytes()
    cfg = synth.SynthConfig(span_mean=500, span_sd=60, seed_base=66000)
    reads = synth.make_reads(10, cfg, genome=genome)
    fe = (_libio.FeRead * len(reads))()
    keep, starts_in, raws = [], [], []
    for i, rd in enumerate(reads):
        clip = (7 + i, 3 + 2 * i)
        ev = synth.to_basecall_events(rd, stay_rate=0.15, clip_bases=clip, seed=i)
        st, starts, bases, read_start = _libio.events_to_starts(ev, True, 4000, 0)
        assert st == _libio.FE_OK
        rec = synth.to_sam_record(rd, soft_clip=clip, seed=i)
        cig, seq = rec['cigar'].encode(), rec['seq'].encode()
        keep += [cig, seq, starts]
        fe[i].cigar, fe[i].flag, fe[i].pos = cig, int(rec['flag']), int(rec['pos'])
        fe[i].seq, fe[i].seq_len = seq, len(seq)
        fe[i].contig, fe[i].contig_len = contig, len(contig)
        fe[i].starts, fe[i].n_starts, fe[i].read_start = starts.ctypes.data, len(starts), read_start
        starts_in.append(starts)
        raws.append(rd.raw)
    R = len(reads)
    cap_s = sum(len(s) for s in starts_in)
    cap_a = 3 * cap_s
    status = np.zeros(R, np.int32); read_start = np.zeros(R, np.int32)
    starts = np.zeros(cap_s, np.int32); starts_off = np.zeros(R + 1, np.int64)
    ra = np.zeros(cap_a, np.uint8); ga = np.zeros(cap_a, np.uint8); align_off = np.zeros(R + 1, np.int64)
    clips = np.zeros(2 * R, np.int64); tallies = np.zeros(4 * R, np.int64)
    rc = _libio.load().nrqio_frontend_batch(fe, R, 4, status.ctypes.data, read_start.ctypes.data, starts.ctypes.data,
                                            cap_s, starts_off.ctypes.data, ra.ctypes.data, ga.ctypes.data, cap_a,
                                            align_off.ctypes.data, clips.ctypes.data, tallies.ctypes.data)
    assert rc == 0 and (status == 0).all()
    # the same reads through the Python objects
    want_reads = []
    for i, rd in enumerate(reads):
        want_reads.append(synth.SynthRead(
            raw=rd.raw, read_start_rel_to_raw=int(read_start[i]),
            starts=starts[starts_off[i]:starts_off[i + 1]].astype(np.int64),
            read_align=ra[align_off[i]:align_off[i + 1]].tobytes(),
            genome_align=ga[align_off[i]:align_off[i + 1]].tobytes()))
        assert want_reads[-1].read_align == rd.read_align and want_reads[-1].genome_align == rd.genome_align
        assert np.array_equal(want_reads[-1].starts, rd.starts)
    want = engine.run(ReadBatch.from_reads(want_reads), ResquiggleParams())
    # ... and the CSR arrays as libnrqio left them, straight into the host entry point
    raw = np.concatenate(raws).astype(np.int16)
    raw_off = np.concatenate([[0], np.cumsum([len(r) for r in raws])]).astype(np.int64)
    A = int(align_off[-1])
    hb = _lib.NrqBatch()
    hb.n_reads = R
    for name, arr in (('raw', raw), ('raw_off', raw_off), ('read_start', read_start), ('starts', starts),
                      ('starts_off', starts_off), ('read_align', ra), ('genome_align', ga), ('align_off', align_off)):
        setattr(hb, name, arr.ctypes.data)
    outs = {'status': np.zeros(R, np.int32), 'n_bases': np.zeros(R, np.int32), 'scale_values': np.zeros(4 * R),
            'new_segs': np.zeros(A + R, np.int32), 'norm_mean': np.zeros(A), 'norm_stdev': np.zeros(A),
            'ev_length': np.zeros(A, np.uint32), 'ev_base': np.zeros(A, np.uint8)}
    hr = _lib.NrqResults()
    for k, v in outs.items():
        setattr(hr, k, v.ctypes.data)
    pc = ResquiggleParams().to_c()
    h = engine.lib.nrq_create(engine.device.index or 0)
    _lib.check(engine.lib.nrq_resquiggle_batch_host(h, C.byref(hb), C.byref(pc), C.byref(hr)), 'host batch')
    engine.lib.nrq_destroy(h)
    assert np.array_equal(outs['status'], want.status) and np.array_equal(outs['n_bases'], want.n_bases)
    for r in range(R):
        o, n = int(align_off[r]), int(want.n_bases[r])
        assert np.array_equal(outs['new_segs'][o + r:o + r + n + 1], want.new_segs_of(r))
        assert np.array_equal(outs['norm_mean'][o:o + n], want.norm_mean[o:o + n])
        assert np.array_equal(outs['ev_length'][o:o + n], want.ev_length[o:o + n])
