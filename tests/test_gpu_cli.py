"""`python -m nanoraw_b200 genome_resquiggle <dir> <genome.fa> --bwa-mem-executable ...` as a user runs it: FAST5
files on disk (this repo's HDF5 reader / writer), alignment and GPU worker processes, a stand-in mapper that
replays SAM records, then every file's RawGenomeCorrected group is compared with the oracle, and the failed-read
table with what the reference reports (resquiggle.py:1116-1198)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from nanoraw_b200 import fast5_io, frontend, synth

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_cli_on_fast5_files_on_disk(tmp_path, monkeypatch):
    from oracle import resquiggle_oracle as orc
    monkeypatch.setattr(fast5_io, '_BACKEND', ['native'])
    cfg = synth.SynthConfig(span_mean=1200, span_sd=200, seed_base=81000)
    genome = synth.make_genome(150_000, seed=31)
    genome_fa = tmp_path / 'genome.fa'
    genome_fa.write_text('>synth_chr1 test genome\n' + genome.tobytes().decode() + '\n')
    reads_dir = tmp_path / 'reads'
    reads_dir.mkdir()
    records, reads, names = {}, [], []
    for i in range(20):
        read = synth.make_read(genome, synth.kmer_levels(cfg), cfg, i)
        events = synth.to_basecall_events(read, stay_rate=0.1, seed=i)
        name = str(reads_dir / ('read_%02d.fast5' % i))
        with fast5_io.open_fast5(name, 'w') as f:
            rd = f.create_group('Raw/Reads/Read_%d' % (i + 1))
            rd.create_dataset('Signal', data=read.raw, compression='gzip')
            rd.attrs['read_id'] = 'read-%d' % i
            rd.attrs['start_time'] = np.uint64(0)
            ch = f.create_group('UniqueGlobalKey/channel_id')
            for k, v in (('offset', 10.0), ('range', 1400.0), ('digitisation', 8192.0), ('sampling_rate', 4000.0)):
                ch.attrs[k] = np.float64(v)
            ch.attrs['channel_number'] = '7'
            f.create_group('Analyses/Basecall_1D_000').attrs['version'] = '1.1.0'
            shifted = events.copy()
            shifted['start'] = events['start'] + read.read_start_rel_to_raw / 4000.0
            f.create_dataset('Analyses/Basecall_1D_000/BaseCalled_template/Events', data=shifted, compression='gzip')
        if i != 5:  # read 5 gets no alignment
            records[('BaseCalled_template:::' + name).replace(' ', '|||')] = synth.to_sam_record(read, seed=i)
        reads.append(read)
        names.append(name)
    (reads_dir / 'not_a_fast5.txt').write_text('hello')  # the reference lists every file of the directory (:1142)
    rec_path = tmp_path / 'records.json'
    rec_path.write_text(json.dumps(records))
    mapper = tmp_path / 'bwa'
    mapper.write_text('#!/bin/sh\nexec %s %s "$@"\n' % (sys.executable, os.path.join(HERE, 'tools', 'fake_bwa.py')))
    mapper.chmod(0o755)
    failed_fn = tmp_path / 'failed.tsv'
    env = dict(os.environ, FAKE_BWA_RECORDS=str(rec_path), PYTHONPATH=ROOT, NRQ_IO_PROCS='2')
    res = subprocess.run(
        [sys.executable, '-m', 'nanoraw_b200', 'genome_resquiggle', str(reads_dir), str(genome_fa),
         '--bwa-mem-executable', str(mapper), '--processes', '2', '--alignment-batch-size', '8',
         '--failed-reads-filename', str(failed_fn)],
        capture_output=True, text=True, timeout=900, env=env, cwd=str(tmp_path))
    assert res.returncode == 0, res.stderr[-3000:]
    assert 'Failed reads summary:' in res.stderr

    from nanoraw_b200 import messages as msg
    failed = dict(ln.split('\t', 1) for ln in failed_fn.read_text().splitlines() if ln)
    assert set(failed) == {msg.NO_ALIGNMENT, msg.CORRUPT_FILE}, failed
    assert failed[msg.NO_ALIGNMENT] == 'BaseCalled_template:::' + names[5]      # :684-685, id format :925
    assert 'not_a_fast5.txt' in failed[msg.CORRUPT_FILE]                          # :966-968

    for i, (name, read) in enumerate(zip(names, reads)):
        f = fast5_io.open_fast5(name, 'r')
        if i == 5:
            assert 'BaseCalled_template' not in f['/Analyses/RawGenomeCorrected_000'].keys()
            continue
        grp = f['/Analyses/RawGenomeCorrected_000']
        assert grp.attrs['nanoraw_version'] == '0.5' and grp.attrs['basecall_group'] == 'Basecall_1D_000'
        sub = grp['BaseCalled_template']
        starts = sub['Alignment/read_segments'][()]
        rows = (sub['Alignment/read_alignment'][()].tobytes().decode(),
                sub['Alignment/genome_alignment'][()].tobytes().decode())
        assert rows[0].encode() == read.read_align and rows[1].encode() == read.genome_align
        read_start = int(sub['Events'].attrs['read_start_rel_to_raw'])
        want = orc.resquiggle_read(read.raw, read_start, starts, rows[0], rows[1])
        ev = sub['Events'][()]
        assert ev.dtype.names == ('norm_mean', 'norm_stdev', 'start', 'length', 'base')
        for col in ('start', 'length', 'base'):
            assert np.array_equal(ev[col], want['events'][col]), (i, col)
        assert np.max(np.abs(ev['norm_mean'] - want['events']['norm_mean'])) < 1e-9
        assert np.max(np.abs(ev['norm_stdev'] - want['events']['norm_stdev'])) < 1e-9
        assert sub.attrs['shift'] == float(want['scale_values'].shift)
        assert sub.attrs['scale'] == float(want['scale_values'].scale)
        assert sub.attrs['norm_type'] == 'median' and sub.attrs['outlier_threshold'] == 5
        assert sub['Alignment'].attrs['mapped_start'] == read.genome_start
        assert sub['Alignment'].attrs['mapped_strand'] == read.strand
        assert sub['Alignment'].attrs['mapped_chrom'] == 'synth_chr1'
