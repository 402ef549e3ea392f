"""hdf5_min.py -- this repo's own HDF5 reader / writer for FAST5 files.

No libhdf5 exists in the image, so the checks are:
* the reader against a file written by the real library (tests/golden/libhdf5_written_v73.mat: a MATLAB v7.3
  file = HDF5 behind a 512-byte user block; it is scipy's own test datum scipy/io/matlab/tests/data/
  testhdf5_7.4_GLNX86.mat, BSD licence), whose expected contents are scipy's documented ones;
* the writer against the format specification through a strict structural walk (alignment, sizes, counts,
  ordering, B-tree keys, heap layout: the things libhdf5 asserts while it opens a file), and against the
  byte layout of the real-library file for the structures both contain;
* write -> read round trips of everything the FAST5 path stores, including in-place modification.
"""
import os
import struct

import numpy as np
import pytest

from nanoraw_b200 import fast5_io, hdf5_min

HERE = os.path.dirname(os.path.abspath(__file__))
REAL = os.path.join(HERE, 'golden', 'libhdf5_written_v73.mat')


@pytest.fixture
def native(monkeypatch):
    monkeypatch.setattr(fast5_io, '_BACKEND', ['native'])


def events_table(n, seed=0):
    rng = np.random.default_rng(seed)
    ev = np.zeros(n, dtype=[('mean', '<f8'), ('start', '<f8'), ('stdv', '<f8'), ('length', '<f8'),
                            ('model_state', 'S5'), ('move', '<i4'), ('weights', '<f4')])
    ev['mean'] = rng.normal(size=n)
    ev['start'] = np.arange(n) * 0.002
    ev['length'] = 0.002
    ev['model_state'] = [b'ACGTA', b'CGTAC'] * (n // 2) + [b'TTTTT'] * (n % 2)
    ev['move'] = rng.integers(0, 3, n)
    return ev


# ------------------------------------------------------------------------------------- reader vs the real library
def test_reads_file_written_by_libhdf5():
    data = open(REAL, 'rb').read()
    root, lossy = hdf5_min.load_tree(data)
    assert lossy == []
    assert root.keys() == ['testdouble']
    ds = root['testdouble']
    assert ds.shape == (9, 1) and ds.dtype == np.dtype('<f8')
    # scipy/io/matlab/tests/test_mio.py: testdouble = linspace(0, 2 pi, 9)
    assert np.allclose(ds[()].ravel(), np.linspace(0, 2 * np.pi, 9), rtol=0, atol=1e-15)
    assert ds.attrs['MATLAB_class'] == b'double'


# ------------------------------------------------------------------------------------- strict structural walk
def _u(d, off, n):
    return int.from_bytes(d[off:off + n], 'little')


def check_structure(d):
    """walk a file the way libhdf5 does when it opens every object, asserting what the library asserts"""
    assert d[:8] == hdf5_min.SIGNATURE
    assert d[8] == 0 and d[13] == 8 and d[14] == 8            # superblock 0, 8-byte offsets and lengths
    leaf_k, int_k = _u(d, 16, 2), _u(d, 18, 2)
    assert _u(d, 24, 8) == 0                                   # base address
    assert _u(d, 40, 8) == len(d)                              # end-of-file address covers exactly the file
    assert _u(d, 56, 8) == 0 and _u(d, 72, 4) == 1             # root entry: name offset 0, cached symbol table
    seen = {'objects': 0, 'chunks': 0, 'attrs': 0, 'groups': 0}

    def header(addr):
        assert addr % 8 == 0 and d[addr] == 1 and d[addr + 1] == 0
        n, size = _u(d, addr + 2, 2), _u(d, addr + 8, 4)
        assert _u(d, addr + 4, 4) == 1                          # reference count
        p, end, msgs = addr + 16, addr + 16 + size, []
        while p < end:
            mtype, msize = _u(d, p, 2), _u(d, p + 2, 2)
            assert msize % 8 == 0 and p + 8 + msize <= end
            msgs.append((mtype, d[p + 8:p + 8 + msize]))
            p += 8 + msize
        assert p == end and len(msgs) == n and end <= len(d)
        return msgs

    def datatype(b, p=0):
        cls, ver, size = b[p] & 15, b[p] >> 4, _u(b, p + 4, 4)
        assert ver == 1 and size > 0
        if cls == 0:
            assert _u(b, p + 8, 2) == 0 and _u(b, p + 10, 2) == 8 * size
            return size, p + 12
        if cls == 1:
            assert _u(b, p + 10, 2) == 8 * size and b[p + 12] + b[p + 13] + 1 == 8 * size
            return size, p + 20
        if cls == 3:
            return size, p + 8
        if cls == 9:
            assert size == 16
            _, q = datatype(b, p + 8)
            return size, q
        assert cls == 6
        q, hi = p + 8, 0
        for _ in range(_u(b, p + 1, 2)):
            e = b.index(b'\x00', q)
            q += (e - q + 8) & ~7
            off = _u(b, q, 4)
            assert b[q + 4] == 0
            msize, q = datatype(b, q + 32)
            assert off >= hi and off + msize <= size                # members in order, inside the record
            hi = off + msize
        return size, q

    def attribute(b):
        assert b[0] == 1
        nsz, tsz, ssz = _u(b, 2, 2), _u(b, 4, 2), _u(b, 6, 2)
        assert b[8 + nsz - 1] == 0
        p = 8 + ((nsz + 7) & ~7)
        esize, used = datatype(b, p)
        assert used - p == tsz
        p += (tsz + 7) & ~7
        assert b[p] == 1
        rank = b[p + 1]
        assert ssz == 8 + 8 * rank * (2 if b[p + 2] & 1 else 1)
        n = 1
        for k in range(rank):
            n *= _u(b, p + 8 + 8 * k, 8)
        p += (ssz + 7) & ~7
        assert len(b) - p >= n * esize and len(b) - p < n * esize + 8
        if b[8 + ((nsz + 7) & ~7)] & 15 == 9:                        # variable-length string elements
            for k in range(n):
                ln, ga, gi = struct.unpack_from('<IQI', b, p + 16 * k)
                assert d[ga:ga + 4] == b'GCOL' and gi >= 1
                q, gend, found = ga + 16, ga + _u(d, ga + 8, 8), False
                while q + 16 <= gend and _u(d, q, 2) != 0:
                    if _u(d, q, 2) == gi:
                        assert _u(d, q + 8, 8) == ln
                        found = True
                    q += 16 + ((_u(d, q + 8, 8) + 7) & ~7)
                assert found
        seen['attrs'] += 1

    def group(btree, heap):
        assert d[heap:heap + 4] == b'HEAP' and d[heap + 4] == 0
        hsize, hfree, hdata = _u(d, heap + 8, 8), _u(d, heap + 16, 8), _u(d, heap + 24, 8)
        assert hsize % 8 == 0 and hdata + hsize <= len(d) and d[hdata:hdata + 8] == b'\x00' * 8
        while hfree != 1:                                            # free list: (next, size) blocks inside the segment
            assert hfree % 8 == 0 and hfree + 16 <= hsize
            nxt, fsz = _u(d, hdata + hfree, 8), _u(d, hdata + hfree + 8, 8)
            assert fsz >= 16 and hfree + fsz <= hsize
            hfree = nxt

        def name_at(off):
            assert off < hsize
            return bytes(d[hdata + off:d.index(b'\x00', hdata + off)])

        names = []

        def node(addr, lo_key):
            assert d[addr:addr + 4] == b'TREE' and d[addr + 4] == 0
            level, n = d[addr + 5], _u(d, addr + 6, 2)
            assert n <= 2 * int_k and addr + 24 + (2 * int_k + 1) * 8 + 2 * int_k * 8 <= len(d)
            p = addr + 24
            assert name_at(_u(d, p, 8)) == lo_key
            for _ in range(n):
                child, key = _u(d, p + 8, 8), name_at(_u(d, p + 16, 8))
                if level:
                    node(child, lo_key)
                else:
                    assert d[child:child + 4] == b'SNOD' and d[child + 4] == 1
                    ns = _u(d, child + 6, 2)
                    assert 1 <= ns <= 2 * leaf_k and child + 8 + 2 * leaf_k * 40 <= len(d)
                    for k in range(ns):
                        e = child + 8 + 40 * k
                        nm = name_at(_u(d, e, 8))
                        assert nm > lo_key and nm <= key                # lo_key < names of this child <= key
                        names.append((nm, _u(d, e + 8, 8), _u(d, e + 16, 4), _u(d, e + 24, 8), _u(d, e + 32, 8)))
                        lo_key = nm
                    assert lo_key == key
                lo_key = key
                p += 16
            return lo_key

        node(btree, b'')
        assert [n[0] for n in names] == sorted(n[0] for n in names) and len({n[0] for n in names}) == len(names)
        seen['groups'] += 1
        for nm, haddr, cache, s0, s1 in names:
            obj(haddr, (s0, s1) if cache == 1 else None)

    def obj(addr, cached):
        seen['objects'] += 1
        msgs = header(addr)
        types = [m[0] for m in msgs]
        for mtype, b in msgs:
            if mtype == 0x0C:
                attribute(b)
        if 0x11 in types:
            b = msgs[types.index(0x11)][1]
            assert cached is None or cached == (_u(b, 0, 8), _u(b, 8, 8))
            group(_u(b, 0, 8), _u(b, 8, 8))
            return
        assert {0x01, 0x03, 0x05, 0x08} <= set(types)
        esize, _ = datatype(msgs[types.index(0x03)][1])
        sp = msgs[types.index(0x01)][1]
        shape = [_u(sp, 8 + 8 * k, 8) for k in range(sp[1])]
        lay = msgs[types.index(0x08)][1]
        assert lay[0] == 3
        if lay[1] == 1:
            a, sz = _u(lay, 2, 8), _u(lay, 10, 8)
            assert sz == int(np.prod(shape)) * esize if shape else sz == esize
            assert a == hdf5_min.UNDEF or a + sz <= len(d)
        else:
            assert lay[1] == 2 and lay[2] == len(shape) + 1 and 0x0B in types
            bt = _u(lay, 3, 8)
            cdims = [_u(lay, 11 + 4 * k, 4) for k in range(len(shape) + 1)]
            assert cdims[-1] == esize and all(c > 0 for c in cdims)
            ksz = 8 + 8 * (len(shape) + 1)
            assert d[bt:bt + 4] == b'TREE' and d[bt + 4] == 1 and d[bt + 5] == 0
            n = _u(d, bt + 6, 2)
            assert bt + 24 + 65 * ksz + 64 * 8 <= len(d)                # a full node of the default K = 32
            p = bt + 24
            for _ in range(n):
                csize, caddr = _u(d, p, 4), _u(d, p + ksz, 8)
                assert csize > 0 and caddr + csize <= len(d)
                assert _u(d, p + 8 + 8 * len(shape), 8) == 0
                p += ksz + 8
                seen['chunks'] += 1

    obj(_u(d, 64, 8), (_u(d, 80, 8), _u(d, 88, 8)))
    return seen


def build_fast5(f, sig, ev):
    rd = f.create_group('Raw/Reads/Read_12')
    rd.attrs['read_id'] = 'abc-def'
    rd.attrs['start_time'] = np.uint64(123456)
    rd.create_dataset('Signal', data=sig, compression='gzip')
    ch = f.create_group('UniqueGlobalKey/channel_id')
    for k, v in (('offset', 3.0), ('range', 1234.5), ('digitisation', 8192.0), ('sampling_rate', 4000.0)):
        ch.attrs[k] = v
    ch.attrs['channel_number'] = b'42'
    bc = f.create_group('Analyses/Basecall_1D_000')
    bc.attrs['version'] = '1.2.3'
    bc.attrs['name'] = np.bytes_(b'albacore')
    bc.create_group('BaseCalled_template').create_dataset('Events', data=ev)


def test_written_file_is_structurally_valid_and_round_trips(tmp_path, native):
    rng = np.random.default_rng(1)
    sig = rng.integers(-500, 1500, size=30000).astype(np.int16)
    ev = events_table(501)
    fn = str(tmp_path / 'read.fast5')
    with fast5_io.open_fast5(fn, 'w') as f:
        build_fast5(f, sig, ev)
        for i in range(300):  # more children than one B-tree node holds: two levels
            f.create_group('Many/g%03d' % i).attrs['i'] = i
        f['Many'].attrs['arr'] = np.arange(5, dtype=np.int32)
        f['Many'].attrs['strs'] = np.array([b'ab', b'cde'])
        f['Many'].attrs['unicode'] = u'café'
        f['Many'].attrs['f32'] = np.float32(2.5)
        f.create_group('Empty')
        f.create_dataset('scalar', data=np.int64(7))
        f.create_dataset('empty', data=np.zeros(0, dtype=np.int16), compression='gzip')
        f.create_dataset('matrix', data=np.arange(12, dtype='<u2').reshape(3, 4), compression='gzip')
    data = open(fn, 'rb').read()
    seen = check_structure(data)
    assert seen['groups'] == 11 + 300 and seen['chunks'] == 2
    with fast5_io.open_fast5(fn, 'r') as f:
        assert np.array_equal(f['Raw/Reads/Read_12/Signal'][()], sig)
        got = f['Analyses/Basecall_1D_000/BaseCalled_template/Events'][()]
        assert got.dtype == ev.dtype and np.array_equal(got, ev)
        raw = list(f['/Raw/Reads/'].values())[0]
        assert raw.attrs['read_id'] == 'abc-def' and raw.attrs['start_time'] == 123456
        assert raw.attrs['start_time'].dtype == np.uint64
        assert f['UniqueGlobalKey/channel_id'].attrs['channel_number'] == b'42'
        assert f['Analyses/Basecall_1D_000'].attrs['version'] == '1.2.3'
        assert sorted(f['Many'].keys()) == ['g%03d' % i for i in range(300)]
        assert all(f['Many/g%03d' % i].attrs['i'] == i for i in (0, 7, 8, 255, 256, 299))
        assert np.array_equal(f['Many'].attrs['arr'], np.arange(5)) and f['Many'].attrs['arr'].dtype == np.int32
        assert list(f['Many'].attrs['strs']) == [b'ab', b'cde']
        assert f['Many'].attrs['unicode'] == u'café' and f['Many'].attrs['f32'] == np.float32(2.5)
        assert f['Empty'].keys() == [] and f['scalar'][()] == 7 and f['empty'].shape == (0,)
        assert np.array_equal(f['matrix'][()], np.arange(12).reshape(3, 4))
        assert 'Analyses/Nope' not in f
        with pytest.raises(KeyError):
            f['Analyses/Nope']


def test_modify_in_place_like_the_resquiggle_writer(tmp_path, native):
    """resquiggle.py:80-135 / 942-988: open r+, drop an old corrected group, add the new one, flush, close"""
    sig = np.arange(-100, 4000, dtype=np.int16)
    ev = events_table(64)
    fn = str(tmp_path / 'read.fast5')
    with fast5_io.open_fast5(fn, 'w') as f:
        build_fast5(f, sig, ev)
    before = os.path.getsize(fn)
    new_ev = np.zeros(20, dtype=[('norm_mean', '<f8'), ('norm_stdev', '<f8'), ('start', '<u4'), ('length', '<u4'),
                                 ('base', 'S1')])
    new_ev['norm_mean'] = np.linspace(-1, 1, 20)
    new_ev['start'] = np.arange(20) * 7
    new_ev['length'] = 7
    new_ev['base'] = list(b'ACGT' * 5)
    new_ev['base'] = np.frombuffer(b'ACGT' * 5, dtype='S1')
    for attempt in range(2):  # the second pass overwrites what the first one wrote (--overwrite)
        f = fast5_io.open_fast5(fn, 'r+')
        if '/Analyses/RawGenomeCorrected_000' in f:
            del f['/Analyses/RawGenomeCorrected_000']
        grp = f['/Analyses'].create_group('RawGenomeCorrected_000')
        grp.attrs['nanoraw_version'] = '0.5'
        grp.attrs['basecall_group'] = 'Basecall_1D_000'
        sub = grp.create_group('BaseCalled_template')
        for k, v in (('shift', 501.5), ('scale', 41.0), ('lower_lim', -5.0), ('upper_lim', 5.0),
                     ('norm_type', 'median'), ('outlier_threshold', 5)):
            sub.attrs[k] = v
        with pytest.raises(TypeError):
            sub.attrs['outlier_threshold'] = None  # h5py has no type for None
        al = sub.create_group('Alignment')
        al.attrs['mapped_start'] = 1234 + attempt
        al.attrs['mapped_strand'] = '+'
        al.create_dataset('read_alignment', data=np.frombuffer(b'ACG-T', dtype='S1'), compression='gzip')
        al.create_dataset('read_segments', data=np.arange(21, dtype=np.int64) * 7, compression='gzip')
        sub.create_dataset('Events', data=new_ev, compression='gzip').attrs['read_start_rel_to_raw'] = 17
        f.flush()
        f.close()
    check_structure(open(fn, 'rb').read())
    assert os.path.getsize(fn) > before
    assert not [n for n in os.listdir(str(tmp_path)) if 'tmp' in n]
    with fast5_io.open_fast5(fn, 'r') as f:
        sub = f['/Analyses/RawGenomeCorrected_000/BaseCalled_template']
        assert sub.attrs['norm_type'] == 'median' and sub.attrs['shift'] == 501.5 and sub.attrs['outlier_threshold'] == 5
        assert sub['Alignment'].attrs['mapped_start'] == 1235
        assert np.array_equal(sub['Events'][()], new_ev) and sub['Events'].attrs['read_start_rel_to_raw'] == 17
        assert sub['Alignment/read_alignment'][()].tobytes() == b'ACG-T'
        assert np.array_equal(f['Raw/Reads/Read_12/Signal'][()], sig)          # untouched content survives
        assert np.array_equal(f['Analyses/Basecall_1D_000/BaseCalled_template/Events'][()], ev)
    with fast5_io.open_fast5(fn, 'r') as f:  # read-only handles never rewrite
        mtime = os.path.getmtime(fn)
    assert os.path.getmtime(fn) == mtime


def test_same_encodings_as_the_real_library(tmp_path, native):
    """byte-for-byte agreement with libhdf5's own output for the messages both files contain: the float64
    datatype, a fixed-string scalar attribute, the symbol-table node, the group B-tree node and the heap header"""
    real = open(REAL, 'rb').read()[512:]
    fn = str(tmp_path / 'm.h5')
    with fast5_io.open_fast5(fn, 'w') as f:
        ds = f.create_dataset('testdouble', data=np.linspace(0, 2 * np.pi, 9).reshape(9, 1))
        ds.attrs['MATLAB_class'] = np.bytes_(b'double')
    mine = open(fn, 'rb').read()
    f64 = bytes.fromhex('11203f000800000000004000340b0034ff030000')
    assert f64 in real and f64 in mine
    attr = bytes.fromhex('01000d00080008004d41544c41425f636c61737300000000')
    assert attr in real and attr in mine
    # the attribute's string datatype: the library wrote null-terminated (MATLAB asked for it), h5py and this
    # writer use null-padded for numpy 'S' data; everything else in the message is identical
    tail_real = real[real.index(attr) + len(attr):][:24]
    tail_mine = mine[mine.index(attr) + len(attr):][:24]
    assert tail_real[0] == 0x13 and tail_mine[0] == 0x13 and (tail_real[1], tail_mine[1]) == (0x00, 0x01)
    assert tail_real[2:] == tail_mine[2:]
    for sig in (b'SNOD', b'TREE', b'HEAP'):
        assert real[real.index(sig):real.index(sig) + 8] == mine[mine.index(sig):mine.index(sig) + 8]
    assert real[8:20] == mine[8:20]  # superblock version bytes, sizes and B-tree ranks (then the status flags)
    root, _ = hdf5_min.load_tree(mine)
    assert np.array_equal(root['testdouble'][()], hdf5_min.load_tree(open(REAL, 'rb').read())[0]['testdouble'][()])


def test_refuses_what_it_cannot_represent(tmp_path, native):
    with pytest.raises(IOError):
        fast5_io.open_fast5(str(tmp_path / 'missing.fast5'), 'r')
    bad = tmp_path / 'bad.fast5'
    bad.write_bytes(b'not an hdf5 file' * 10)
    with pytest.raises(IOError):
        fast5_io.open_fast5(str(bad), 'r')
    f = fast5_io.open_fast5(str(tmp_path / 'x.fast5'), 'w')
    f.create_dataset('v', data=np.arange(3))
    f['v'].attrs['big'] = np.zeros(10000)  # a version-1 object header message holds at most 64 KB
    with pytest.raises(hdf5_min.Unsupported):
        f.flush()
    assert not os.path.exists(str(tmp_path / 'x.fast5'))  # nothing half-written is left behind
    del f['v'].attrs['big']
    f.close()
    assert fast5_io.open_fast5(str(tmp_path / 'x.fast5'), 'r')['v'][()].tolist() == [0, 1, 2]


def test_reader_handles_new_style_structures():
    """superblock 2, version-2 object headers with creation-order tracking and a continuation block, link messages
    (compact group), dataspace v2, attribute v3, compact and single-chunk (layout v4) datasets: none of which the
    writer emits.  The bytes are laid out here by hand from the format specification (checksums are not verified
    by the reader and are left zero)."""
    import zlib
    buf = bytearray(48)  # superblock v2 goes in last

    def put(b):
        while len(buf) % 8:
            buf.append(0)
        a = len(buf)
        buf.extend(b)
        return a

    def msg(mtype, body, track=False):
        return struct.pack('<BHB', mtype, len(body), 0) + (b'\x00\x00' if track else b'') + body

    def ohdr(msgs, flags=0):
        body = b''.join(msgs)
        head = b'OHDR' + bytes([2, flags])
        if flags & 0x20:
            head += b'\x00' * 16
        if flags & 0x10:
            head += struct.pack('<HH', 8, 6)
        size_field = {0: '<B', 1: '<H', 2: '<I'}[flags & 3]
        return head + struct.pack(size_field, len(body)) + body + b'\x00' * 4

    f64 = struct.pack('<BBBBIHHBBBBI', 0x11, 0x20, 63, 0, 8, 0, 64, 52, 11, 0, 52, 1023)
    i16 = struct.pack('<BBBBIHH', 0x10, 0x08, 0, 0, 2, 0, 16)
    space1 = struct.pack('<BBBBQ', 2, 1, 0, 1, 5)            # dataspace v2, simple, rank 1, 5 elements
    scalar = struct.pack('<BBBB', 2, 0, 0, 0)                # dataspace v2, scalar
    # attribute v3: name 'gain' (UTF-8), float64 scalar
    attr = struct.pack('<BBHHHB', 3, 0, 5, len(f64), len(scalar), 1) + b'gain\x00' + f64 + scalar + struct.pack('<d', 2.5)
    # dataset A: compact layout (v3), int16[5]
    vals = np.array([3, -1, 4, -1, 5], dtype='<i2')
    compact = struct.pack('<BBH', 3, 0, vals.nbytes) + vals.tobytes()
    ds_a = put(ohdr([msg(0x01, space1), msg(0x03, i16), msg(0x08, compact), msg(0x0C, attr)]))
    # dataset B: layout v4, single chunk, deflate, float64[5]; second half of its messages in a continuation block
    data = np.linspace(0, 1, 5)
    comp = zlib.compress(data.tobytes(), 4)
    chunk = put(comp)
    filt = struct.pack('<BBHHHI', 2, 1, 1, 1, 1, 4)          # pipeline v2: deflate, flags 1, one client value
    lay4 = struct.pack('<BBBBB', 4, 2, 0x02, 2, 1) + bytes([5, 8]) + bytes([1]) + struct.pack('<QI', len(comp), 0) + \
        struct.pack('<Q', chunk)
    cont_body = msg(0x0B, filt, True) + msg(0x08, lay4, True)
    cont = put(b'OCHK' + cont_body + b'\x00' * 4)
    ds_b = put(ohdr([msg(0x01, space1, True), msg(0x03, f64, True),
                     msg(0x10, struct.pack('<QQ', cont, len(cont_body) + 8), True)], flags=0x04 | 0x20))
    # a sub-group holding B, and the root group holding A and the sub-group: link messages with 1-byte name length
    def link(name, addr):
        return struct.pack('<BB', 1, 0) + bytes([len(name)]) + name + struct.pack('<Q', addr)
    linfo = struct.pack('<BBQQ', 0, 0, hdf5_min.UNDEF, hdf5_min.UNDEF)
    sub = put(ohdr([msg(0x02, linfo), msg(0x06, link(b'B', ds_b))]))
    root = put(ohdr([msg(0x02, linfo), msg(0x06, link(b'A', ds_a)), msg(0x06, link(b'sub', sub))], flags=1))
    buf[0:48] = hdf5_min.SIGNATURE + struct.pack('<BBBBQQQQI', 2, 8, 8, 0, 0, hdf5_min.UNDEF, len(buf), root, 0)
    tree, lossy = hdf5_min.load_tree(bytes(buf))
    assert lossy == [] and sorted(tree.keys()) == ['A', 'sub']
    assert tree['A'].dtype == np.dtype('<i2') and np.array_equal(tree['A'][()], vals)
    assert tree['A'].attrs['gain'] == 2.5
    assert np.array_equal(tree['sub/B'][()], data) and tree['sub/B'].compression == 'gzip'
    # and what was read can be written back in the conservative format and read again
    again, _ = hdf5_min.load_tree(hdf5_min.dump_tree(tree))
    check_structure(hdf5_min.dump_tree(tree))
    assert np.array_equal(again['A'][()], vals) and np.array_equal(again['sub/B'][()], data)
    assert again['A'].attrs['gain'] == 2.5


def _random_tree(rng, depth=0):
    """a random group: attributes, datasets of assorted types (some gzip), sub-groups"""
    dtypes = ['<i1', '<u1', '<i2', '<u2', '<i4', '<u4', '<i8', '<u8', '<f4', '<f8', 'S1', 'S7',
              [('a', '<f8'), ('b', '<u4'), ('c', 'S3')], [('x', '<i2'), ('y', '<f4')]]
    node = {'attrs': {}, 'datasets': {}, 'groups': {}}
    for k in range(int(rng.integers(0, 5))):
        kind = int(rng.integers(0, 6))
        name = 'attr_%d_%s' % (k, 'x' * int(rng.integers(0, 9)))
        if kind == 0:
            node['attrs'][name] = int(rng.integers(-2**40, 2**40))
        elif kind == 1:
            node['attrs'][name] = float(rng.normal())
        elif kind == 2:
            node['attrs'][name] = ''.join(chr(int(c)) for c in rng.integers(0x20, 0x24f, int(rng.integers(0, 20))))
        elif kind == 3:
            node['attrs'][name] = np.bytes_(bytes(rng.integers(65, 91, int(rng.integers(1, 12))).astype(np.uint8)))
        elif kind == 4:
            node['attrs'][name] = rng.integers(0, 100, int(rng.integers(1, 6))).astype(rng.choice(['<i4', '<u2', '<f8']))
        else:
            node['attrs'][name] = np.float32(rng.normal())
    for k in range(int(rng.integers(0, 4))):
        dt = np.dtype(dtypes[int(rng.integers(0, len(dtypes)))])
        shape = tuple(int(v) for v in rng.integers(0, 40, int(rng.integers(1, 3))))
        raw = rng.integers(0, 256, int(np.prod(shape)) * dt.itemsize, dtype=np.uint8).tobytes()
        arr = np.frombuffer(raw, dtype=dt).reshape(shape).copy()
        node['datasets']['ds_%d' % k] = (arr, bool(rng.integers(0, 2)))
    if depth < 3:
        for k in range(int(rng.integers(0, 4 if depth else 12))):
            node['groups']['grp_%02d' % k] = _random_tree(rng, depth + 1)
    return node


def _write_tree(h5, node):
    for k, v in node['attrs'].items():
        h5.attrs[k] = v
    for k, (arr, gz) in node['datasets'].items():
        h5.create_dataset(k, data=arr, compression='gzip' if gz else None)
    for k, sub in node['groups'].items():
        _write_tree(h5.create_group(k), sub)


def _same(a, b):
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        a, b = np.asarray(a), np.asarray(b)
        return a.dtype == b.dtype and a.shape == b.shape and a.tobytes() == b.tobytes()
    return type(a) == type(b) and (a == b or (a != a and b != b)) if not isinstance(a, (str, bytes)) else a == b


def _check_tree(h5, node):
    assert sorted(h5.attrs.keys()) == sorted(node['attrs'])
    for k, v in node['attrs'].items():
        got = h5.attrs[k]
        want = np.int64(v) if isinstance(v, int) else np.float64(v) if isinstance(v, float) else v
        assert _same(got, want), (k, got, want)
    assert sorted(h5.keys()) == sorted(list(node['datasets']) + list(node['groups']))
    for k, (arr, gz) in node['datasets'].items():
        got = h5[k][()]
        assert got.dtype == arr.dtype and got.shape == arr.shape and got.tobytes() == arr.tobytes(), k
    for k, sub in node['groups'].items():
        _check_tree(h5[k], sub)


@pytest.mark.parametrize('seed', range(12))
def test_random_trees_round_trip(tmp_path, native, seed):
    """random object trees (nested groups up to 12 wide, assorted attribute and dataset types, empty and gzip
    datasets) survive write -> structural walk -> read -> modify -> read"""
    rng = np.random.default_rng(7000 + seed)
    tree = _random_tree(rng)
    fn = str(tmp_path / 'rand.h5')
    with fast5_io.open_fast5(fn, 'w') as f:
        _write_tree(f, tree)
    check_structure(open(fn, 'rb').read())
    with fast5_io.open_fast5(fn, 'r') as f:
        _check_tree(f, tree)
    extra = _random_tree(rng, depth=2)
    with fast5_io.open_fast5(fn, 'r+') as f:
        _write_tree(f.create_group('added_later'), extra)
    tree['groups']['added_later'] = extra
    check_structure(open(fn, 'rb').read())
    with fast5_io.open_fast5(fn, 'r') as f:
        _check_tree(f, tree)
