"""libnrqio's FAST5 functions (include/nrq_io.h: nrqio_f5_*) on the CPU: the raw-signal reader and the
RawGenomeCorrected appender against this repo's Python HDF5 reader / writer (hdf5_min.py), which is itself pinned on
a libhdf5-written file and a strict structural walk (tests/test_hdf5_min.py).  resquiggle.py:327-347 is what is read,
:80-135 the layout that is written."""
import os
import shutil
import struct
import zlib

import numpy as np
import pytest

from nanoraw_b200 import _libio, fast5_io, hdf5_min, synth
from nanoraw_b200 import batch as nbatch
from nanoraw_b200 import nanoraw_helper as nh
from nanoraw_b200 import resquiggle as rsq
from nanoraw_b200.frontend import genomeLoc, readInfo

from test_hdf5_min import check_structure


@pytest.fixture
def native(monkeypatch):
    monkeypatch.setattr(fast5_io, '_BACKEND', ['native'])


@pytest.fixture(scope='module')
def reads():
    return synth.make_reads(6, synth.config_short())


def _files(tmp_path, reads):
    paths = []
    for i, r in enumerate(reads):
        fn = str(tmp_path / ('read_%d.fast5' % i))
        synth.write_basecalled_fast5_job((fn, r))
        paths.append(fn)
    return paths


def _read_all(paths, threads=2):
    b = _libio.F5RawBatch(paths, threads=threads)
    off = np.zeros(len(paths), np.int64)
    off[1:] = np.cumsum(b.n_samples)[:-1]
    dst = np.full(int(b.n_samples.sum()) + 8, -12345, np.int16)
    status = b.read_into(dst, off)
    b.close()
    return b, off, dst, status


# ----------------------------------------------------------------------------------------------- reader
@pytest.mark.parametrize('threads', [1, 4])
def test_raw_signals_equal_the_python_reader(tmp_path, native, reads, threads):
    paths = _files(tmp_path, reads)
    b, off, dst, status = _read_all(paths, threads)
    assert (b.status == _libio.F5_OK).all() and (status == _libio.F5_OK).all()
    for i, r in enumerate(reads):
        chan, raw, _, _ = rsq._read_raw_inputs(paths[i], 'median', 'Basecall_1D_000', 'BaseCalled_template')
        assert b.n_samples[i] == len(raw)
        assert np.array_equal(dst[off[i]:off[i] + len(raw)], raw) and np.array_equal(raw, r.raw)
        assert tuple(b.channel[i]) == (chan.offset, chan.range, chan.digitisation, float(chan.sampling_rate))
    assert (dst[int(b.n_samples.sum()):] == -12345).all()  # nothing written past the last signal


def test_contiguous_and_empty_signals(tmp_path, native):
    rng = np.random.default_rng(5)
    for name, raw, comp in (('plain', rng.integers(-500, 900, 3000).astype(np.int16), None),
                            ('empty', np.zeros(0, np.int16), 'gzip'),
                            ('one', np.array([7], np.int16), 'gzip')):
        fn = str(tmp_path / (name + '.fast5'))
        f = fast5_io.open_fast5(fn, 'w')
        f.create_group('Raw/Reads/Read_9').create_dataset('Signal', data=raw, compression=comp)
        ch = f.create_group('UniqueGlobalKey/channel_id')
        ch.attrs['offset'] = np.int32(-3)          # integer attributes are converted like float ones
        ch.attrs['range'] = np.float32(1.5)
        ch.attrs['digitisation'] = np.uint16(8192)
        ch.attrs['sampling_rate'] = 4000
        ch.attrs['channel_number'] = b'12'
        f.close()
        b, off, dst, status = _read_all([fn])
        assert b.status[0] == 0 and status[0] == 0 and b.n_samples[0] == len(raw)
        assert np.array_equal(dst[:len(raw)], raw)
        assert tuple(b.channel[0]) == (-3.0, 1.5, 8192.0, 4000.0)


def test_first_read_by_name_and_missing_pieces(tmp_path, native):
    fn = str(tmp_path / 'two.fast5')
    f = fast5_io.open_fast5(fn, 'w')
    f.create_group('Raw/Reads/Read_20').create_dataset('Signal', data=np.arange(10, dtype=np.int16))
    f.create_group('Raw/Reads/Read_10').create_dataset('Signal', data=np.arange(5, dtype=np.int16), compression='gzip')
    ch = f.create_group('UniqueGlobalKey/channel_id')
    for k in ('offset', 'range', 'digitisation', 'sampling_rate'):
        ch.attrs[k] = 1.0
    ch.attrs['channel_number'] = '3'
    f.close()
    b, off, dst, status = _read_all([fn])
    assert b.status[0] == 0 and b.n_samples[0] == 5          # Read_10 sorts first, as in h5py's listing
    # no channel_id group (mux scan file, nanoraw_helper.py:157-160), no raw read, not HDF5, no file
    cases = {}
    f = fast5_io.open_fast5(str(tmp_path / 'nochan.fast5'), 'w')
    f.create_group('Raw/Reads/Read_1').create_dataset('Signal', data=np.arange(4, dtype=np.int16))
    f.close()
    cases['nochan.fast5'] = _libio.F5_MISSING
    f = fast5_io.open_fast5(str(tmp_path / 'noraw.fast5'), 'w')
    f.create_group('Analyses')
    f.close()
    cases['noraw.fast5'] = _libio.F5_MISSING
    (tmp_path / 'text.fast5').write_bytes(b'not an hdf5 file' * 100)
    cases['text.fast5'] = _libio.F5_FORMAT
    (tmp_path / 'empty.fast5').write_bytes(b'')
    cases['empty.fast5'] = _libio.F5_FORMAT
    cases['absent.fast5'] = _libio.F5_IO
    f = fast5_io.open_fast5(str(tmp_path / 'float.fast5'), 'w')      # a signal that is not int16: the Python reader converts
    f.create_group('Raw/Reads/Read_1').create_dataset('Signal', data=np.arange(4, dtype=np.float32))
    f.close()
    cases['float.fast5'] = _libio.F5_MISSING  # (channel_id is looked up after the signal; the type is refused first)
    names = sorted(cases)
    b = _libio.F5RawBatch([str(tmp_path / n) for n in names] + [fn], threads=3)
    got = dict(zip(names, b.status[:-1]))
    got_float = got.pop('float.fast5')
    cases.pop('float.fast5')
    assert got_float == _libio.F5_UNSUPPORTED
    assert got == cases and b.status[-1] == 0
    dst = np.zeros(8, np.int16)
    off = np.full(len(names) + 1, -1, np.int64)
    off[-1] = 2
    st = b.read_into(dst, off)
    assert st[-1] == 0 and np.array_equal(dst[2:7], np.arange(5)) and dst[7] == 0
    b.close()
    # truncated files: every prefix is refused or read, never a crash
    data = open(fn, 'rb').read()
    for cut in list(range(0, len(data), 97)) + [len(data) - 1]:
        (tmp_path / 'cut.fast5').write_bytes(data[:cut])
        b = _libio.F5RawBatch([str(tmp_path / 'cut.fast5')])
        if b.status[0] == 0:
            st = b.read_into(np.zeros(16, np.int16), [0])
            assert st[0] in (_libio.F5_OK, _libio.F5_FORMAT)
        else:
            assert b.status[0] in (_libio.F5_FORMAT, _libio.F5_MISSING, _libio.F5_UNSUPPORTED)
        b.close()


def _new_style_fast5(raw, chunk, shuffle=True, fletcher=True, mask_last=False):
    """a FAST5 file in the forms libhdf5 writes with libver='latest' plus a chunked, filtered signal: superblock 2,
    version-2 object headers, link messages, attribute v3, layout v3 over a v1 chunk B-tree with several chunks,
    shuffle + deflate + fletcher32 (bytes laid out by hand from the format specification; checksums zero)"""
    buf = bytearray(48)

    def put(b):
        while len(buf) % 8:
            buf.append(0)
        a = len(buf)
        buf.extend(b)
        return a

    def msg(mtype, body):
        return struct.pack('<BHB', mtype, len(body), 0) + body

    def ohdr(msgs):
        body = b''.join(msgs)
        return b'OHDR' + bytes([2, 2]) + struct.pack('<I', len(body)) + body + b'\x00' * 4

    f64 = struct.pack('<BBBBIHHBBBBI', 0x11, 0x20, 63, 0, 8, 0, 64, 52, 11, 0, 52, 1023)
    i16 = struct.pack('<BBBBIHH', 0x10, 0x08, 0, 0, 2, 0, 16)
    s2 = struct.pack('<BBBBI', 0x13, 0x01, 0, 0, 2)
    scalar = struct.pack('<BBBB', 2, 0, 0, 0)

    def attr(name, dt, data):
        nm = name.encode() + b'\x00'
        return msg(0x0C, struct.pack('<BBHHHB', 3, 0, len(nm), len(dt), len(scalar), 0) + nm + dt + scalar + data)

    def link(name, addr):
        return msg(0x06, struct.pack('<BB', 1, 0) + bytes([len(name)]) + name + struct.pack('<Q', addr))

    linfo = msg(0x02, struct.pack('<BBQQ', 0, 0, hdf5_min.UNDEF, hdf5_min.UNDEF))
    n = len(raw)
    keys = []
    for lo in range(0, n, chunk):
        block = np.zeros(chunk, dtype='<i2')
        block[:min(chunk, n - lo)] = raw[lo:lo + chunk]
        data = block.tobytes()
        last = lo + chunk >= n
        mask = 0
        if shuffle:
            data = np.frombuffer(data, np.uint8).reshape(-1, 2).T.tobytes()
        if mask_last and last:
            mask |= 2 if shuffle else 1  # the deflate stage was skipped for this chunk
        else:
            data = zlib.compress(data, 6)
        if fletcher:
            data += b'\x00\x00\x00\x00'
        keys.append((len(data), mask, lo, put(data)))
    node = bytearray(b'TREE' + struct.pack('<BBH', 1, 0, len(keys)) + struct.pack('<QQ', hdf5_min.UNDEF, hdf5_min.UNDEF))
    for size, mask, lo, addr in keys:
        node += struct.pack('<IIQQ', size, mask, lo, 0) + struct.pack('<Q', addr)
    node += struct.pack('<IIQQ', 0, 0, ((n + chunk - 1) // chunk) * chunk, 0)
    btree = put(bytes(node))
    filters = []
    if shuffle:
        filters.append(struct.pack('<HHHI', 2, 0, 1, 2))
    filters.append(struct.pack('<HHHI', 1, 1, 1, 6))
    if fletcher:
        filters.append(struct.pack('<HHH', 3, 0, 0))
    pipeline = msg(0x0B, struct.pack('<BB', 2, len(filters)) + b''.join(filters))
    layout = msg(0x08, struct.pack('<BBBQII', 3, 2, 2, btree, chunk, 2))
    space = msg(0x01, struct.pack('<BBBBQQ', 2, 1, 1, 1, n, n))
    sig = put(ohdr([space, msg(0x03, i16), pipeline, layout]))
    rd = put(ohdr([linfo, link(b'Signal', sig), attr('start_time', struct.pack('<BBBBIHH', 0x10, 0, 0, 0, 8, 0, 64),
                                                      struct.pack('<Q', 0))]))
    reads_g = put(ohdr([linfo, link(b'Read_7', rd)]))
    raw_g = put(ohdr([linfo, link(b'Reads', reads_g)]))
    chan = put(ohdr([linfo, attr('offset', f64, struct.pack('<d', 4.0)), attr('range', f64, struct.pack('<d', 1467.6)),
                     attr('digitisation', f64, struct.pack('<d', 8192.0)),
                     attr('sampling_rate', f64, struct.pack('<d', 4000.0)), attr('channel_number', s2, b'42')]))
    ugk = put(ohdr([linfo, link(b'channel_id', chan)]))
    root = put(ohdr([linfo, link(b'Raw', raw_g), link(b'UniqueGlobalKey', ugk)]))
    buf[0:48] = hdf5_min.SIGNATURE + struct.pack('<BBBBQQQQI', 2, 8, 8, 0, 0, hdf5_min.UNDEF, len(buf), root, 0)
    return bytes(buf)


@pytest.mark.parametrize('shuffle,fletcher,mask_last', [(True, True, False), (False, False, False), (True, False, True),
                                                        (False, True, True)])
def test_chunked_filtered_signal_in_new_style_structures(tmp_path, shuffle, fletcher, mask_last):
    rng = np.random.default_rng(11)
    for n, chunk in ((10_000, 4096), (4096, 4096), (5, 16), (8193, 1024)):
        raw = np.cumsum(rng.integers(-40, 41, n)).astype(np.int16)
        data = _new_style_fast5(raw, chunk, shuffle, fletcher, mask_last)
        tree, _ = hdf5_min.load_tree(data)
        assert np.array_equal(tree['Raw/Reads/Read_7/Signal'][()], raw)   # the Python reader agrees on the file
        fn = str(tmp_path / 'new_style.fast5')
        open(fn, 'wb').write(data)
        b, off, dst, status = _read_all([fn])
        assert b.status[0] == 0 and status[0] == 0 and b.n_samples[0] == n
        assert np.array_equal(dst[:n], raw) and (dst[n:] == -12345).all()
        assert tuple(b.channel[0]) == (4.0, 1467.6, 8192.0, 4000.0)


# ----------------------------------------------------------------------------------------------- writer
def _tree(fn):
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
    return out, lossy


def _corrected(read, seed):
    rng = np.random.default_rng(seed)
    nb = len(read.genome_align.replace(b'-', b''))
    ev = np.zeros(nb, dtype=np.dtype(rsq.EVENT_DTYPE))
    ev['norm_mean'] = rng.normal(size=nb)
    ev['norm_stdev'] = rng.random(nb)
    ev['length'] = rng.integers(4, 30, nb)
    ev['start'] = np.cumsum(ev['length']) - ev['length']
    ev['base'] = np.frombuffer(read.genome_align.replace(b'-', b''), dtype='S1')
    ra, ga = np.frombuffer(read.read_align, dtype='S1'), np.frombuffer(read.genome_align, dtype='S1')
    segs = np.asarray(read.starts, dtype=np.int64)
    chunks = [_libio.deflate(x.tobytes()) for x in (ev, ra, ga, segs)]
    return (nb, len(ra), len(segs)) + tuple(chunks)


def _native_write(fn, packed, loc, info, sv, read_start, subgroup='BaseCalled_template', norm_type='median', thresh=5):
    chunks = packed[3:]
    buf = np.frombuffer(b''.join(chunks), dtype=np.uint8)
    ln = np.array([len(c) for c in chunks], np.int64)
    co = np.zeros(4, np.int64)
    co[1:] = np.cumsum(ln)[:-1]
    return int(_libio.f5_write_corrected(
        [fn], 'RawGenomeCorrected_000', subgroup, norm_type, thresh, [loc.Chrom], [loc.Strand],
        np.array([sv.shift, sv.scale, sv.lower_lim, sv.upper_lim]),
        np.array([loc.Start, info.ClipStart, info.ClipEnd, info.Insertions, info.Deletions, info.Matches,
                  info.Mismatches, read_start]), [packed[0]], [packed[1]], [packed[2]], buf, co, ln, threads=1)[0])


def test_appended_group_equals_the_python_writers(tmp_path, native, reads):
    paths = _files(tmp_path, reads[:3])
    for i, fn in enumerate(paths):
        fa, fb = fn + '.python', fn + '.native'
        shutil.copy(fn, fa)
        shutil.copy(fn, fb)
        packed = _corrected(reads[i], i)
        loc = genomeLoc(1000 + i, '+-'[i & 1], 'chr%d_é' % i)
        info = readInfo('x', 'BaseCalled_template', 1 + i, 2, 3, 4, 5, 6)
        sv = nh.scaleValues(431.5 + i, 57.25, -5.0, 5.0)
        thresh = 5 if i != 1 else 4.5
        rsq.write_packed_fast5_group(fa, loc, info, 77 + i, packed, sv, 'RawGenomeCorrected_000',
                                     'BaseCalled_template', 'median', thresh)
        assert _native_write(fb, packed, loc, info, sv, 77 + i, thresh=thresh) == _libio.F5_OK
        ta, la = _tree(fa)
        tb, lb = _tree(fb)
        assert la == [] and lb == []
        assert ta.keys() == tb.keys()
        for k in ta:
            assert ta[k] == tb[k], k
        check_structure(open(fb, 'rb').read())           # what libhdf5 asserts while it opens every object
        # the group is there now: h5py's create_group would raise (-> the WRITE_BACK error of :126-127)
        assert _native_write(fb, packed, loc, info, sv, 0) == _libio.F5_EXISTS
        assert _tree(fb)[0] == tb
        # a second subgroup next to it (2D reads: template and complement, :1064-1075)
        assert _native_write(fb, packed, loc, info, sv, 5, subgroup='BaseCalled_complement') == _libio.F5_OK
        rsq.write_packed_fast5_group(fa, loc, info, 5, packed, sv, 'RawGenomeCorrected_000', 'BaseCalled_complement',
                                     'median', 5)
        assert _tree(fa)[0] == _tree(fb)[0]
        check_structure(open(fb, 'rb').read())
        # the Python writer can still rewrite what was appended, and the raw signal is untouched
        f = fast5_io.open_fast5(fb, 'r+')
        del f['/Analyses/RawGenomeCorrected_000/BaseCalled_template']
        f.close()
        b, off, dst, status = _read_all([fb])
        assert status[0] == 0 and np.array_equal(dst[:len(reads[i].raw)], reads[i].raw)


def test_append_next_to_many_siblings(tmp_path, native, reads):
    """more members than one symbol table node holds (2 * leaf K = 8), then more than one B-tree node's worth"""
    fn = _files(tmp_path, reads[:1])[0]
    packed = _corrected(reads[0], 0)
    loc, info, sv = genomeLoc(5, '+', 'chr1'), readInfo('x', 'BaseCalled_template', 0, 0, 0, 0, 0, 0), \
        nh.scaleValues(1.0, 2.0, -5.0, 5.0)
    names = ['BaseCalled_%03d' % k for k in range(40)]
    for k, name in enumerate(names):
        assert _native_write(fn, packed, loc, info, sv, k, subgroup=name) == _libio.F5_OK
        if k in (0, 7, 8, 9, 39):
            check_structure(open(fn, 'rb').read())
    f = fast5_io.open_fast5(fn, 'r')
    grp = f['/Analyses/RawGenomeCorrected_000']
    assert sorted(grp.keys()) == names
    for k, name in enumerate(names):
        assert grp[name + '/Events'].attrs['read_start_rel_to_raw'] == k
        assert len(grp[name + '/Events'][()]) == packed[0]
    assert grp.attrs['basecall_group'] == 'Basecall_1D_000'   # the group's own attributes are untouched
    f.close()


def test_files_outside_the_slice_are_left_alone(tmp_path, native, reads):
    packed = _corrected(reads[0], 0)
    loc, info, sv = genomeLoc(5, '+', 'chr1'), readInfo('x', 'BaseCalled_template', 0, 0, 0, 0, 0, 0), \
        nh.scaleValues(1.0, 2.0, -5.0, 5.0)
    fn = str(tmp_path / 'new_style.fast5')                     # superblock 2, link-message groups, no /Analyses
    data = _new_style_fast5(np.arange(100, dtype=np.int16), 64)
    open(fn, 'wb').write(data)
    assert _native_write(fn, packed, loc, info, sv, 0) == _libio.F5_UNSUPPORTED
    assert open(fn, 'rb').read() == data
    fn = _files(tmp_path, reads[:1])[0]
    f = fast5_io.open_fast5(fn, 'r+')
    del f['/Analyses/RawGenomeCorrected_000']                  # prep_fast5 has not run
    f.close()
    data = open(fn, 'rb').read()
    assert _native_write(fn, packed, loc, info, sv, 0) == _libio.F5_MISSING
    assert open(fn, 'rb').read() == data
    assert _native_write(str(tmp_path / 'absent.fast5'), packed, loc, info, sv, 0) == _libio.F5_IO


def test_h5py_reads_the_appended_group(tmp_path, native, reads):
    h5py = pytest.importorskip('h5py')
    fn = _files(tmp_path, reads[:1])[0]
    packed = _corrected(reads[0], 0)
    loc, info, sv = genomeLoc(5, '+', 'chr1'), readInfo('x', 'BaseCalled_template', 0, 0, 0, 0, 0, 0), \
        nh.scaleValues(1.0, 2.0, -5.0, 5.0)
    assert _native_write(fn, packed, loc, info, sv, 9) == _libio.F5_OK
    with h5py.File(fn, 'r') as f:
        g = f['/Analyses/RawGenomeCorrected_000/BaseCalled_template']
        assert g.attrs['norm_type'] == 'median' and g['Alignment'].attrs['mapped_chrom'] == 'chr1'
        assert len(g['Events']) == packed[0] and g['Events'].attrs['read_start_rel_to_raw'] == 9
        assert g['Alignment/read_segments'][()].tolist() == list(reads[0].starts)


# ----------------------------------------------------------------------------------------------- pipeline glue
class _FakeResults(object):
    pass


def test_batch_loader_fills_the_csr_arrays(tmp_path, native, reads):
    """BatchPipeline._load_native: libnrqio inflates into the batcher's raw array; a file it refuses is read by the
    Python reader; a file nobody can read fails with the reference's message (:343-347)"""
    from nanoraw_b200 import messages as msg
    paths = _files(tmp_path, reads)
    # file 1: float32 signal (outside the native slice, the Python reader converts); file 4: not a FAST5 file
    f = fast5_io.open_fast5(paths[1], 'r+')
    del f['/Raw/Reads/Read_1/Signal']
    f['/Raw/Reads/Read_1'].create_dataset('Signal', data=reads[1].raw.astype(np.float32), compression='gzip')
    f.close()
    open(paths[4], 'wb').write(b'garbage')
    pend = []
    for fn, r in zip(paths, reads):
        info = readInfo(os.path.basename(fn), 'BaseCalled_template', 0, 0, 0, 0, 0, 0)
        loc = genomeLoc(int(r.genome_start), r.strand, r.chrom)
        pend.append(rsq._PendingRead(fn, ((r.read_align, r.genome_align), loc, r.starts,
                                          int(r.read_start_rel_to_raw), info)))
    for norm_type in ('median', 'pA_raw'):
        pipe = rsq.BatchPipeline(norm_type, 5, None, None, 'Basecall_1D_000', 'RawGenomeCorrected_000', True, None)
        assert pipe._native(pend)
        b = nbatch.PinnedBatcher(max_reads=2, raw_capacity=1 << 12, pinned=False)   # has to grow
        lb = pipe._load_native(pend, b)[0]
        assert [pr.fast5_fn for pr in lb.ready] == [p for i, p in enumerate(paths) if i != 4]
        assert [(e, pr.fast5_fn) for e, pr in lb.failed] == [(msg.OPEN_FOR_RESQUIGGLE, paths[4])]
        assert lb.unreadable == {} and b.n_reads == 5
        for k, i in enumerate([0, 1, 2, 3, 5]):
            r = reads[i]
            assert np.array_equal(b.raw[b.raw_off[k]:b.raw_off[k + 1]], r.raw)
            assert np.array_equal(b.starts[b.starts_off[k]:b.starts_off[k + 1]], r.starts)
            assert b.read_align[b.align_off[k]:b.align_off[k + 1]].tobytes() == r.read_align
            assert b.genome_align[b.align_off[k]:b.align_off[k + 1]].tobytes() == r.genome_align
            assert b.read_start[k] == r.read_start_rel_to_raw
        if norm_type == 'pA_raw':
            assert b.has_given
            sv = b.scale_values_in[:20].reshape(5, 4)
            assert np.allclose(sv[:, 0], -10.0) and np.allclose(sv[:, 1], 8192.0 / 1400.0) and np.isnan(sv[:, 2:]).all()
        else:
            assert not b.has_given
    assert not rsq.BatchPipeline('pA', 5, None, None, 'a', 'b', True, None)._native(pend)


def test_batch_writer_falls_back_per_read(tmp_path, native, reads):
    """BatchPipeline._write_native: appended by libnrqio where it can, the Python writer otherwise, the reference's
    messages for the failures (:82-86, :126-127)"""
    from nanoraw_b200 import messages as msg
    paths = _files(tmp_path, reads[:5])
    packs = [_corrected(r, i) for i, r in enumerate(reads[:5])]
    buf = np.frombuffer(b''.join(b''.join(p[3:]) for p in packs), dtype=np.uint8)
    ln = np.array([[len(c) for c in p[3:]] for p in packs], np.int64)
    off = (np.cumsum(ln.ravel()) - ln.ravel()).reshape(-1, 4)
    writers = []
    for i, (fn, r) in enumerate(zip(paths, reads)):
        info = readInfo(os.path.basename(fn), 'BaseCalled_template', 0, 1, 2, 3, 4, 5)
        loc = genomeLoc(int(r.genome_start), r.strand, r.chrom)
        writers.append(rsq._PendingRead(fn, ((r.read_align, r.genome_align), loc, r.starts, 11 * i, info)))
    # read 1: already corrected (EXISTS); read 2: file vanished (IO); read 3: a numpy int32 attribute (Python types it)
    assert _native_write(paths[1], packs[1], writers[1].genome_loc, writers[1].read_info,
                         nh.scaleValues(1.0, 2.0, -5.0, 5.0), 0) == 0
    os.remove(paths[2])
    odd = np.array([0, 0, 0, 1, 0], dtype=np.int32)
    ints = np.array([(pr.genome_loc.Start, 0, 1, 2, 3, 4, 5, pr.read_start) for pr in writers], dtype=np.int64)
    sv = np.tile(np.array([400.0, 50.0, -5.0, 5.0]), (5, 1))
    job = dict(writers=writers, sv=sv, ints=ints, odd=odd, n_bases=np.array([p[0] for p in packs], np.int32),
               n_align=np.array([p[1] for p in packs], np.int64), n_segs=np.array([p[2] for p in packs], np.int64),
               chunks=buf, off=off, ln=ln)
    pipe = rsq.BatchPipeline('median', 5, None, None, 'Basecall_1D_000', 'RawGenomeCorrected_000', True, None)
    errs = pipe._write_native(job)
    assert errs == [None, msg.WRITE_BACK, msg.OPEN_FOR_WRITE, None, None]
    for i in (0, 3, 4):
        f = fast5_io.open_fast5(paths[i], 'r')
        ev = f['/Analyses/RawGenomeCorrected_000/BaseCalled_template/Events']
        assert ev.attrs['read_start_rel_to_raw'] == 11 * i and len(ev[()]) == packs[i][0]
        assert f['/Analyses/RawGenomeCorrected_000/BaseCalled_template'].attrs['shift'] == 400.0
        f.close()
