"""Front end -> queue -> GPU re-squiggle worker -> FAST5 write, in one process: basecalled FAST5 files
(in-memory backend) + a stand-in mapper that replays SAM records -> RawGenomeCorrected groups, checked
against the oracle run on the inputs the front end produced."""
import json
import os
import queue
import sys

import numpy as np
import pytest

from nanoraw_b200 import fast5_io, synth

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
FAKE_BWA = os.path.join(HERE, 'tools', 'fake_bwa.py')


def make_basecalled_fast5(name, read, events, start_time=0):
    f = fast5_io.open_fast5(name, 'w')
    rd = f.create_group('Raw/Reads/Read_3')
    rd.create_dataset('Signal', data=read.raw)
    rd.attrs['read_id'] = read.read_id
    rd.attrs['start_time'] = np.uint64(start_time)
    ch = f.create_group('UniqueGlobalKey/channel_id')
    for k, v in (('offset', 10.0), ('range', 1400.0), ('digitisation', 8192.0), ('sampling_rate', 4000.0)):
        ch.attrs[k] = np.float64(v)
    ch.attrs['channel_number'] = '11'
    grp = f.create_group('Analyses/Basecall_1D_000')
    grp.attrs['version'] = '1.1.0'
    f.create_group('Analyses/Basecall_1D_000/BaseCalled_template').create_dataset('Events', data=events)
    f.close()


@pytest.mark.parametrize('backend', ['mem', 'native'])
def test_alignment_worker_to_resquiggle_worker(tmp_path, monkeypatch, backend):
    """backend 'mem': FAST5 files in the in-process store; 'native': real files on disk, read and rewritten
    by this repo's own HDF5 reader / writer (hdf5_min.py)"""
    from nanoraw_b200 import frontend
    monkeypatch.setattr(fast5_io, '_BACKEND', [backend if backend == 'native' else 'auto'])
    prefix = 'mem://pipeline/' if backend == 'mem' else str(tmp_path) + '/'
    from nanoraw_b200 import resquiggle as rsq
    from oracle import resquiggle_oracle as orc

    cfg = synth.SynthConfig(span_mean=900, span_sd=150, seed_base=71000)
    genome = synth.make_genome(120_000, seed=23)
    genome_fa = tmp_path / 'genome.fa'
    genome_fa.write_text('>synth_chr1 test genome\n' + genome.tobytes().decode() + '\n')
    n_reads = 12
    records, names, reads = {}, [], []
    for i in range(n_reads):
        read = synth.make_read(genome, synth.kmer_levels(cfg), cfg, i)
        # the basecaller saw a few extra bases each side that the mapper soft-clips
        clip = (i % 3, (2 * i) % 5)
        events = synth.to_basecall_events(read, stay_rate=0.12, clip_bases=clip, seed=i)
        name = prefix + 'read_%02d.fast5' % i
        # samples taken by the clipped leading bases (a base starts at every moving event)
        base_events = np.flatnonzero(events['move'] == 1)
        lead = int(np.round(events['start'][base_events[clip[0]]] * 4000.0))
        # the clipped bases sit just before the read's own signal inside the raw array
        shifted = events.copy()
        shifted['start'] = events['start'] + (read.read_start_rel_to_raw - lead) / 4000.0
        assert read.read_start_rel_to_raw - lead >= 0
        make_basecalled_fast5(name, read, shifted)
        rec = synth.to_sam_record(read, soft_clip=clip, seed=i)
        records[('BaseCalled_template:::' + name).replace(' ', '|||')] = rec
        names.append(name)
        reads.append(read)
    # one file with no alignment, one file without basecalls
    unaligned = prefix + 'unaligned.fast5'
    extra = synth.make_read(genome, synth.kmer_levels(cfg), cfg, 99)
    make_basecalled_fast5(unaligned, extra, synth.to_basecall_events(extra, seed=99))
    rec_path = tmp_path / 'records.json'
    rec_path.write_text(json.dumps(records))
    monkeypatch.setenv('FAKE_BWA_RECORDS', str(rec_path))
    mapper = tmp_path / 'bwa'
    mapper.write_text('#!/bin/sh\nexec %s %s "$@"\n' % (sys.executable, FAKE_BWA))
    mapper.chmod(0o755)

    fast5_q, basecalls_q, failed_q = queue.Queue(), queue.Queue(), queue.Queue()
    fast5_q.put(names[:7] + [unaligned])
    fast5_q.put(names[7:])
    frontend.alignment_worker(fast5_q, basecalls_q, failed_q, str(genome_fa), str(mapper), 'bwa_mem',
                              'Basecall_1D_000', ['BaseCalled_template'], 'RawGenomeCorrected_000', False, 1)
    items = []
    while not basecalls_q.empty():
        items.append(basecalls_q.get())
    assert len(items) == len(names)
    for it in items:
        basecalls_q.put(it)
    basecalls_q.put((None, None))
    rsq.resquiggle_worker(basecalls_q, failed_q, 'Basecall_1D_000', 'RawGenomeCorrected_000', 'median', 5,
                          None, None, True, None)
    failures = {}
    while not failed_q.empty():
        err, rid = failed_q.get()
        failures[rid] = err
    assert failures == {'BaseCalled_template:::' + unaligned:
                        'Alignment not produced. Potentially failed to locate BWA index files.'}

    by_name = dict(items)
    for name, read in zip(names, reads):
        (rows, loc, starts, read_start, info), = by_name[name]
        assert rows[0].encode() == read.read_align and rows[1].encode() == read.genome_align
        assert loc.Strand == read.strand and loc.Start == read.genome_start
        want = orc.resquiggle_read(read.raw, int(read_start), starts, rows[0], rows[1])
        sub = fast5_io.open_fast5(name, 'r')['/Analyses/RawGenomeCorrected_000/BaseCalled_template']
        ev = sub['Events'][()]
        assert np.array_equal(ev['start'], want['events']['start'])
        assert np.array_equal(ev['length'], want['events']['length'])
        assert np.array_equal(ev['base'], want['events']['base'])
        assert np.max(np.abs(ev['norm_mean'] - want['events']['norm_mean'])) < 1e-9
        assert sub.attrs['shift'] == float(want['scale_values'].shift)
        assert sub.attrs['scale'] == float(want['scale_values'].scale)
        assert sub['Alignment'].attrs['mapped_start'] == read.genome_start
        assert sub['Alignment'].attrs['clipped_bases_start'] == info.ClipStart
        assert np.array_equal(sub['Alignment/read_segments'][()], starts)
        assert sub['Events'].attrs['read_start_rel_to_raw'] == read_start
    # second run without --overwrite: every file is skipped (resume semantics of prep_fast5)
    fast5_q.put(names)
    frontend.alignment_worker(fast5_q, basecalls_q, failed_q, str(genome_fa), str(mapper), 'bwa_mem',
                              'Basecall_1D_000', ['BaseCalled_template'], 'RawGenomeCorrected_000', False, 1)
    assert basecalls_q.empty()
    skipped = [failed_q.get() for _ in range(failed_q.qsize())]
    assert len(skipped) == len(names)
    assert all(e == 'Raw genome corrected data exists for this read and --overwrite is not set.'
               for e, _ in skipped)


def test_native_file_path_equals_the_python_one(tmp_path, monkeypatch):
    """BatchPipeline on files on disk, once with libnrqio's FAST5 functions and once with hdf5_min.py on the I/O
    processes (NRQ_FAST5_NATIVE=0): the same failures with the same messages, and object trees that are equal
    file by file -- including the files the native path hands to the Python reader / writer."""
    import shutil
    from nanoraw_b200 import hdf5_min, messages as msg
    from nanoraw_b200 import resquiggle as rsq
    from nanoraw_b200.frontend import genomeLoc, readInfo
    monkeypatch.setattr(fast5_io, '_BACKEND', ['native'])
    monkeypatch.setenv('NRQ_IO_PROCS', '0')
    reads = synth.make_reads(9, synth.config_short())
    dirs = {}
    for mode in ('native', 'python'):
        d = tmp_path / mode
        d.mkdir()
        dirs[mode] = d
    base = []
    for i, r in enumerate(reads):
        fn = str(dirs['native'] / ('read_%d.fast5' % i))
        synth.write_basecalled_fast5_job((fn, r))
        base.append(fn)
    # 2: float32 signal (libnrqio refuses the type, the Python reader converts); 3: not an HDF5 file;
    # 4: already corrected (create_group raises); 5: no RawGenomeCorrected_000 group (prep_fast5 has not run)
    f = fast5_io.open_fast5(base[2], 'r+')
    del f['/Raw/Reads/Read_1/Signal']
    f['/Raw/Reads/Read_1'].create_dataset('Signal', data=reads[2].raw.astype(np.float32), compression='gzip')
    f.close()
    open(base[3], 'wb').write(b'\x89HDF\r\n\x1a\n' + b'\x00' * 50)
    f = fast5_io.open_fast5(base[4], 'r+')
    f['/Analyses/RawGenomeCorrected_000'].create_group('BaseCalled_template')
    f.close()
    f = fast5_io.open_fast5(base[5], 'r+')
    del f['/Analyses/RawGenomeCorrected_000']
    f.close()
    for fn in base:
        shutil.copy(fn, str(dirs['python'] / os.path.basename(fn)))

    def run(mode):
        monkeypatch.setenv('NRQ_FAST5_NATIVE', '1' if mode == 'native' else '0')
        pend = []
        for i, r in enumerate(reads):
            fn = str(dirs[mode] / ('read_%d.fast5' % i))
            info = readInfo(os.path.basename(fn), 'BaseCalled_template', 1, 2, 3, 4, 5, 6)
            # 6: a numpy int32 where an int is usual (typed by h5py's rules on the Python writer)
            start = np.int32(r.genome_start) if i == 6 else int(r.genome_start)
            loc = genomeLoc(start, r.strand, r.chrom)
            pend.append(rsq._PendingRead(fn, ((r.read_align, r.genome_align), loc, r.starts,
                                              int(r.read_start_rel_to_raw), info)))
        pipe = rsq.BatchPipeline('median', 5, None, None, 'Basecall_1D_000', 'RawGenomeCorrected_000', True, None)
        assert pipe._native(pend) == (mode == 'native')
        failed = pipe.feed(pend[:5]) + pipe.feed(pend[5:]) + pipe.drain()
        return sorted((os.path.basename(pr.fast5_fn), text) for text, pr in failed)

    def tree(fn):
        root, lossy = hdf5_min.load_tree(open(fn, 'rb').read())
        out = {}

        def visit(g, path):
            out[path + '@'] = {k: (np.asarray(v).dtype.str, np.asarray(v).tolist()) for k, v in g.attrs.items()}
            if isinstance(g, fast5_io.MemGroup):
                for k, c in g._children.items():
                    visit(c, path + '/' + k)
            else:
                a = g[()]
                out[path] = (str(a.dtype), a.shape, a.tobytes(), g.compression)
        visit(root, '')
        return out

    failed_native, failed_python = run('native'), run('python')
    assert failed_native == failed_python
    assert failed_native == [('read_3.fast5', msg.OPEN_FOR_RESQUIGGLE), ('read_4.fast5', msg.WRITE_BACK),
                             ('read_5.fast5', msg.WRITE_BACK)]
    for i in range(len(reads)):
        if i == 3:
            continue
        a = tree(str(dirs['native'] / ('read_%d.fast5' % i)))
        b = tree(str(dirs['python'] / ('read_%d.fast5' % i)))
        assert a.keys() == b.keys(), i
        for k in a:
            assert a[k] == b[k], (i, k)
        if i not in (4, 5):
            assert '/Analyses/RawGenomeCorrected_000/BaseCalled_template/Events' in a
    assert a['/Analyses/RawGenomeCorrected_000/BaseCalled_template/Alignment@']['mapped_start'][0] == '<i8'
    six = tree(str(dirs['native'] / 'read_6.fast5'))
    assert six['/Analyses/RawGenomeCorrected_000/BaseCalled_template/Alignment@']['mapped_start'][0] == '<i4'
