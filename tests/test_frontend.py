"""Front end (FAST5 events / SAM record -> hot-path inputs) against vectors computed by the
REAL reference (tests/golden/frontend_v1.json, made by tests/golden/make_frontend_golden.py)."""
import json
import os

import numpy as np
import pytest

from nanoraw_b200 import fast5_io, frontend, synth

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, 'golden', 'frontend_v1.json')))
GENOME = synth.make_genome(GOLD['genome_len'], seed=GOLD['genome_seed'])
GENOME_INDEX = {'synth_chr1': GENOME.tobytes().decode()}


def _err(e):
    return '%s: %s' % (type(e).__name__, e)


@pytest.mark.parametrize('entry', GOLD['sam'], ids=[e['record']['cigar'][:24] for e in GOLD['sam']])
def test_parse_sam_record(entry):
    if 'error' in entry:
        with pytest.raises(Exception) as ei:
            frontend.parse_sam_record(entry['record'], GENOME_INDEX)
        assert _err(ei.value) == entry['error']
        return
    rows, loc, cs, ce = frontend.parse_sam_record(entry['record'], GENOME_INDEX)
    assert rows[0] == entry['read_row'] and rows[1] == entry['genome_row']
    assert [loc.Start, loc.Strand, loc.Chrom] == entry['loc']
    assert [cs, ce] == entry['clip']
    assert frontend.as_pairs(rows)[0] == (entry['read_row'][0], entry['genome_row'][0])


def _events_file(entry):
    cfg = synth.SynthConfig(span_mean=300, span_sd=20, seed_base=900)
    read = synth.make_read(GENOME, synth.kmer_levels(cfg), cfg, entry['index'])
    spec = entry['spec']
    ev = synth.to_basecall_events(read, stay_rate=spec.get('stay_rate', 0.1),
                                  lead_stays=spec.get('lead_stays', 0),
                                  trail_stays=spec.get('trail_stays', 0), seed=entry['seed'])
    if spec.get('zero_len'):
        ev['length'][5] = 0.0
    if spec.get('all_stays'):
        ev['move'][1:] = 0
    name = 'mem://test_frontend/%d' % entry['index']
    f = fast5_io.MemFile(name, 'w')
    rd = f.create_group('Raw/Reads/Read_7')
    rd.create_dataset('Signal', data=read.raw)
    rd.attrs['read_id'] = 'golden-read'
    rd.attrs['start_time'] = np.uint64(entry['start_time'])
    ch = f.create_group('UniqueGlobalKey/channel_id')
    for k, v in (('offset', 10.0), ('range', 1400.0), ('digitisation', 8192.0),
                 ('sampling_rate', 4000.0)):
        ch.attrs[k] = np.float64(v)
    ch.attrs['channel_number'] = '7'
    grp = f.create_group('Analyses/Basecall_1D_000')
    if entry['version'] is not None:
        grp.attrs['version'] = entry['version']
    f.create_group('Analyses/Basecall_1D_000/BaseCalled_template').create_dataset('Events', data=ev)
    f.close()
    return name, read


@pytest.mark.parametrize('entry', GOLD['events'], ids=[json.dumps(e['spec']) for e in GOLD['events']])
@pytest.mark.filterwarnings('ignore:overflow encountered')
def test_get_read_data(entry):
    name, read = _events_file(entry)
    if 'error' in entry:
        with pytest.raises(Exception) as ei:
            frontend.get_read_data(name, 'Basecall_1D_000', 'BaseCalled_template')
        assert _err(ei.value) == entry['error']
        return
    rs, starts, calls, chan, rid = frontend.get_read_data(name, 'Basecall_1D_000', 'BaseCalled_template')
    assert int(rs) == entry['read_start']
    assert [int(v) for v in starts] == entry['starts']
    assert ''.join(calls) == entry['basecalls']
    assert int(chan.sampling_rate) == entry['sampling_rate'] and rid == entry['read_id']
    if not entry['spec'].get('lead_stays') and not entry['spec'].get('trail_stays'):
        # stays removed: the basecaller segmentation of the synthetic read comes back (a stay that
        # ends the read is clipped off, so only the last boundary may differ)
        assert len(starts) == len(read.starts)
        assert np.array_equal(starts[:-1], read.starts[:-1]) and starts[-1] <= read.starts[-1]


@pytest.mark.parametrize('entry', GOLD['clip'], ids=[str(e['clip']) for e in GOLD['clip']])
def test_fix_raw_starts_for_clipped_bases(entry):
    s, rs = frontend.fix_raw_starts_for_clipped_bases(
        entry['clip'][0], entry['clip'][1], np.array(entry['starts_in']), 100)
    assert [int(v) for v in s] == entry['starts'] and int(rs) == entry['read_start']


def test_sam_round_trip_and_tallies():
    cfg = synth.SynthConfig(span_mean=500, span_sd=30, seed_base=950)
    for i in range(6):
        read = synth.make_read(GENOME, synth.kmer_levels(cfg), cfg, i)
        rec = synth.to_sam_record(read, soft_clip=(i, 2 * i), seed=i)
        rows, loc, cs, ce = frontend.parse_sam_record(rec, GENOME_INDEX)
        assert rows[0].encode() == read.read_align and rows[1].encode() == read.genome_align
        assert (cs, ce) == (i, 2 * i) and loc.Strand == read.strand and loc.Start == read.genome_start
        assert frontend.alignment_tallies(rows) == read.tallies()


def test_best_mapq_record_wins_and_missing_reads_fail():
    cfg = synth.SynthConfig(span_mean=200, span_sd=10, seed_base=960)
    read = synth.make_read(GENOME, synth.kmer_levels(cfg), cfg, 0)
    name = 'BaseCalled_template:::some file.fast5'
    good = synth.to_sam_record(read, mapq=60)
    good['qName'] = name.replace(' ', '|||')
    worse = dict(good, mapq='3', pos='17')
    lines = ['@SQ\tSN:synth_chr1\tLN:%d' % len(GENOME),
             '\t'.join(worse[k] for k in frontend.SAM_FIELDS),
             '\t'.join(good[k] for k in frontend.SAM_FIELDS)]
    reads = {name: None, 'BaseCalled_template:::other.fast5': None}
    failed, parsed = frontend.parse_sam_output(lines, reads, GENOME_INDEX)
    assert parsed[name][1].Start == read.genome_start
    assert failed == [('Alignment not produced. Potentially failed to locate BWA index files.',
                       'BaseCalled_template:::other.fast5')]


def test_prep_fast5_resume_semantics():
    name = 'mem://prep/a.fast5'
    f = fast5_io.MemFile(name, 'w')
    f.create_group('Analyses/Basecall_1D_000')
    f.close()
    assert frontend.prep_fast5(name, 'Basecall_1D_000', 'RawGenomeCorrected_000', False, True) is None
    # second run without --overwrite: skipped (this is how an interrupted run resumes)
    err = frontend.prep_fast5(name, 'Basecall_1D_000', 'RawGenomeCorrected_000', False, True)
    assert err == ('Raw genome corrected data exists for this read and --overwrite is not set.', name)
    assert frontend.prep_fast5(name, 'Basecall_1D_000', 'RawGenomeCorrected_000', True, True) is None
    attrs = fast5_io.MemFile(name)['/Analyses/RawGenomeCorrected_000'].attrs
    assert attrs['nanoraw_version'] == '0.5' and attrs['basecall_group'] == 'Basecall_1D_000'
    assert frontend.prep_fast5(name, 'Nope', 'X', False, True)[0].startswith('FAST5 basecall-group')
    assert frontend.prep_fast5('mem://prep/missing', 'a', 'b', False, True)[0] == \
        'Error opening file. Likely a corrupted file.'
