"""The native front end (libnrqio: nrqio_events_to_starts, nrqio_sam_to_rows, nrqio_clip_starts,
nrqio_frontend_batch) against the vectors computed by the REAL reference (tests/golden/frontend_v1.json) and
against the Python mirror of resquiggle.py:447-460, 583-660, 767-907 on random cases."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from nanoraw_b200 import _libio, frontend, messages as msg, synth

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, 'golden', 'frontend_v1.json')))
GENOME = synth.make_genome(GOLD['genome_len'], seed=GOLD['genome_seed'])
CONTIG = GENOME.tobytes()

MESSAGES = {_libio.FE_BAD_CIGAR: msg.BAD_CIGAR, _libio.FE_SEQ_CIGAR_DISAGREE: msg.SEQ_CIGAR_DISAGREE,
            _libio.FE_ALL_STAYS: msg.ALL_STAYS, _libio.FE_EVENTS_BEFORE_RAW: msg.EVENTS_BEFORE_RAW,
            _libio.FE_TOO_FEW_SEGMENTS: msg.TOO_FEW_SEGMENTS, _libio.FE_ZERO_LENGTH: msg.ZERO_LENGTH_INPUT}


def test_status_messages_are_the_reference_strings():
    for code, text in MESSAGES.items():
        assert _libio.fe_message(code) == text


@pytest.mark.parametrize('entry', GOLD['sam'], ids=[e['record']['cigar'][:24] for e in GOLD['sam']])
def test_sam_to_rows_against_the_reference(entry):
    rec = entry['record']
    st, rr, gr, clip, tallies = _libio.sam_to_rows(rec['cigar'], int(rec['flag']), int(rec['pos']), rec['seq'], CONTIG)
    if 'error' in entry:
        assert st != _libio.FE_OK
        if st in MESSAGES:  # (an AssertionError in the reference carries the same text)
            assert MESSAGES[st] in entry['error']
        return
    assert st == _libio.FE_OK
    assert rr.decode() == entry['read_row'] and gr.decode() == entry['genome_row']
    assert list(clip) == entry['clip']
    assert tuple(tallies) == frontend.alignment_tallies((entry['read_row'], entry['genome_row']))


def _events_for(entry):
    cfg = synth.SynthConfig(span_mean=300, span_sd=20, seed_base=900)
    read = synth.make_read(GENOME, synth.kmer_levels(cfg), cfg, entry['index'])
    spec = entry['spec']
    ev = synth.to_basecall_events(read, stay_rate=spec.get('stay_rate', 0.1), lead_stays=spec.get('lead_stays', 0),
                                  trail_stays=spec.get('trail_stays', 0), seed=entry['seed'])
    if spec.get('zero_len'):
        ev['length'][5] = 0.0
    if spec.get('all_stays'):
        ev['move'][1:] = 0
    return ev


@pytest.mark.parametrize('entry', GOLD['events'], ids=[json.dumps(e['spec']) + str(e['version']) for e in GOLD['events']])
def test_events_to_starts_against_the_reference(entry):
    ev = _events_for(entry)
    after = frontend._albacore_after_1_0(entry['version'] if entry['version'] is not None else '0.0')
    st, starts, bases, read_start = _libio.events_to_starts(ev, after, 4000, entry['start_time'])
    if 'error' in entry:
        assert st != _libio.FE_OK and MESSAGES[st] in entry['error']
        return
    assert st == _libio.FE_OK
    assert starts.tolist() == entry['starts'] and bases.decode() == entry['basecalls']
    assert read_start == entry['read_start']


@pytest.mark.parametrize('entry', GOLD['clip'])
def test_clip_starts_against_the_reference(entry):
    starts = np.array(entry['starts_in'], dtype=np.int64)
    n = C.c_int64(len(starts))
    rs = C.c_int64(100 if 'read_start_in' not in entry else entry['read_start_in'])
    want_s, want_rs = frontend.fix_raw_starts_for_clipped_bases(entry['clip'][0], entry['clip'][1],
                                                                np.array(entry['starts_in']), rs.value)
    st = _libio.load().nrqio_clip_starts(starts.ctypes.data, C.byref(n), C.byref(rs), entry['clip'][0], entry['clip'][1])
    assert st == 0 and starts[:n.value].tolist() == entry['starts'] == list(want_s)
    assert rs.value == want_rs


def _random_record(rng, k):
    """a SAM record with soft / hard clips, leading and trailing insertions / deletions, either strand"""
    ops = []
    if rng.random() < 0.3:
        ops.append((int(rng.integers(1, 9)), 'H'))
    if rng.random() < 0.5:
        ops.append((int(rng.integers(1, 30)), 'S'))
    if rng.random() < 0.2:
        ops.append((int(rng.integers(1, 4)), rng.choice(['I', 'D'])))
    for _ in range(int(rng.integers(1, 12))):
        ops.append((int(rng.integers(1, 40)), rng.choice(['M', '=', 'X', 'M', 'M'])))
        if rng.random() < 0.8:
            ops.append((int(rng.integers(1, 5)), rng.choice(['I', 'D', 'N', 'P'])))
    ops.append((int(rng.integers(1, 40)), 'M'))
    if rng.random() < 0.2:
        ops.append((int(rng.integers(1, 4)), rng.choice(['I', 'D'])))
    if rng.random() < 0.5:
        ops.append((int(rng.integers(1, 30)), 'S'))
    if rng.random() < 0.3:
        ops.append((int(rng.integers(1, 9)), 'H'))
    n_read = sum(n for n, op in ops if op in 'MIP=XS')
    seq = ''.join(rng.choice(list('ACGTN'), n_read, p=[.24, .24, .24, .24, .04]))
    if rng.random() < 0.1:
        seq = seq[:-2]  # disagrees with the CIGAR
    return {'qName': 'r%d' % k, 'flag': str(int(rng.choice([0, 16, 2048, 2064]))), 'rName': 'synth_chr1',
            'pos': str(int(rng.integers(1, len(CONTIG) - 2000))), 'mapq': '60',
            'cigar': ''.join('%d%s' % o for o in ops) if rng.random() > 0.03 else '*', 'seq': seq}


def test_native_front_end_equals_the_python_mirror_on_random_records():
    rng = np.random.default_rng(2026)
    index = {'synth_chr1': CONTIG.decode()}
    n_ok = 0
    for k in range(400):
        rec = _random_record(rng, k)
        st, rr, gr, clip, tallies = _libio.sam_to_rows(rec['cigar'], int(rec['flag']), int(rec['pos']), rec['seq'], CONTIG)
        try:
            rows, loc, cs, ce = frontend.parse_sam_record(rec, index)
        except Exception as e:  # noqa: BLE001
            assert st != _libio.FE_OK, (rec, str(e))
            if st in MESSAGES:
                assert MESSAGES[st] == str(e)
            continue
        assert st == _libio.FE_OK, (rec, _libio.fe_message(st))
        n_ok += 1
        assert (rr.decode(), gr.decode()) == rows and clip == (cs, ce)
        assert tuple(tallies) == frontend.alignment_tallies(rows)
    assert n_ok > 150


@pytest.mark.parametrize('dtypes', [('<f4', '<f4'), ('<f8', '<f8'), ('<f8', '<f4')])
@pytest.mark.parametrize('after', [True, False])
def test_events_to_starts_equals_the_python_mirror(dtypes, after):
    """both Albacore generations, float32 and float64 columns (the product is taken in the column's precision)"""
    from nanoraw_b200 import fast5_io
    rng = np.random.default_rng(7)
    cfg = synth.SynthConfig(span_mean=250, span_sd=40, seed_base=950)
    for k in range(12):
        read = synth.make_read(GENOME, synth.kmer_levels(cfg), cfg, k)
        ev0 = synth.to_basecall_events(read, stay_rate=0.25, lead_stays=int(rng.integers(0, 3)),
                                       trail_stays=int(rng.integers(0, 3)), seed=k)
        dt = [(n, dtypes[0] if n == 'start' else dtypes[1] if n == 'length' else ev0.dtype[n]) for n in ev0.dtype.names]
        ev = np.empty(len(ev0), dtype=dt)
        for n in ev0.dtype.names:
            ev[n] = ev0[n]
        ev['start'] += np.asarray(0.25 * k, dtype=ev['start'].dtype)
        t0 = int(round(float(ev['start'][0]) * 4000)) - int(rng.integers(0, 3))
        name = 'mem://fe_native/%s_%d_%d' % (dtypes[0] + dtypes[1], after, k)
        f = fast5_io.MemFile(name, 'w')
        rd = f.create_group('Raw/Reads/Read_7')
        rd.create_dataset('Signal', data=read.raw)
        rd.attrs['read_id'] = 'x'
        rd.attrs['start_time'] = np.uint64(max(t0, 0))
        ch = f.create_group('UniqueGlobalKey/channel_id')
        for kk, v in (('offset', 10.0), ('range', 1400.0), ('digitisation', 8192.0), ('sampling_rate', 4000.0)):
            ch.attrs[kk] = np.float64(v)
        ch.attrs['channel_number'] = '7'
        f.create_group('Analyses/Basecall_1D_000').attrs['version'] = '1.1.0' if after else '0.9.1'
        f.create_group('Analyses/Basecall_1D_000/BaseCalled_template').create_dataset('Events', data=ev)
        want = frontend.get_read_data_py(name, 'Basecall_1D_000', 'BaseCalled_template')
        st, starts, bases, read_start = _libio.events_to_starts(ev, after, 4000, max(t0, 0))
        assert st == _libio.FE_OK
        assert starts.tolist() == [int(v) for v in want[1]] and bases.decode() == ''.join(want[2])
        assert read_start == want[0]


def test_frontend_batch_fills_the_csr_arrays():
    """nrqio_frontend_batch: SAM records + stay-free starts in, nrq_batch CSR arrays out == the per-read Python path"""
    rng = np.random.default_rng(11)
    index = {'synth_chr1': CONTIG.decode()}
    recs, starts_in, keep = [], [], []
    for k in range(60):
        rec = _random_record(rng, k)
        n_read = len(rec['seq'])
        st_arr = np.concatenate([[0], np.cumsum(rng.integers(2, 20, n_read))]).astype(np.int64)
        recs.append(rec)
        starts_in.append(st_arr)
    arr = (_libio.FeRead * len(recs))()
    for i, (rec, st_arr) in enumerate(zip(recs, starts_in)):
        cig, seq = rec['cigar'].encode(), rec['seq'].encode()
        keep += [cig, seq]
        arr[i].cigar, arr[i].flag, arr[i].pos = cig, int(rec['flag']), int(rec['pos'])
        arr[i].seq, arr[i].seq_len = seq, len(seq)
        arr[i].contig, arr[i].contig_len = CONTIG, len(CONTIG)
        arr[i].starts, arr[i].n_starts, arr[i].read_start = st_arr.ctypes.data, len(st_arr), 500 + i
    R = len(recs)
    cap_s = sum(len(s) for s in starts_in)
    cap_a = 4 * cap_s
    status = np.zeros(R, np.int32); read_start = np.zeros(R, np.int32)
    starts = np.zeros(cap_s, np.int32); starts_off = np.zeros(R + 1, np.int64)
    ra = np.zeros(cap_a, np.uint8); ga = np.zeros(cap_a, np.uint8); align_off = np.zeros(R + 1, np.int64)
    clip = np.zeros(2 * R, np.int64); tallies = np.zeros(4 * R, np.int64)
    for threads in (1, 4):
        rc = _libio.load().nrqio_frontend_batch(arr, R, threads, status.ctypes.data, read_start.ctypes.data,
                                                starts.ctypes.data, cap_s, starts_off.ctypes.data, ra.ctypes.data,
                                                ga.ctypes.data, cap_a, align_off.ctypes.data, clip.ctypes.data,
                                                tallies.ctypes.data)
        assert rc == 0
        n_ok = 0
        for i, (rec, st_arr) in enumerate(zip(recs, starts_in)):
            try:
                rows, loc, cs, ce = frontend.parse_sam_record(rec, index)
                if cs >= len(st_arr):
                    raise IndexError
                want_s, want_rs = frontend.fix_raw_starts_for_clipped_bases(cs, ce, st_arr, 500 + i)
            except Exception:  # noqa: BLE001
                assert status[i] != 0 and starts_off[i + 1] == starts_off[i] and align_off[i + 1] == align_off[i]
                continue
            n_ok += 1
            assert status[i] == 0
            assert ra[align_off[i]:align_off[i + 1]].tobytes().decode() == rows[0]
            assert ga[align_off[i]:align_off[i + 1]].tobytes().decode() == rows[1]
            assert starts[starts_off[i]:starts_off[i + 1]].tolist() == [int(v) for v in want_s]
            assert read_start[i] == want_rs and (clip[2 * i], clip[2 * i + 1]) == (cs, ce)
            assert tuple(tallies[4 * i:4 * i + 4]) == frontend.alignment_tallies(rows)
        assert n_ok > 30
    assert _libio.load().nrqio_frontend_batch(arr, R, 1, status.ctypes.data, read_start.ctypes.data,
                                              starts.ctypes.data, 3, starts_off.ctypes.data, ra.ctypes.data,
                                              ga.ctypes.data, cap_a, align_off.ctypes.data, clip.ctypes.data,
                                              tallies.ctypes.data) == _libio.FE_CAPACITY
