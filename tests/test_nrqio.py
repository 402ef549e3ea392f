"""libnrqio (include/nrq_io.h) on the CPU: the deflate encoder against zlib's inflate, the batch packing against
numpy, and the packed FAST5 write-back against the array-by-array one (resquiggle.py:80-135 is the layout)."""
import ctypes as C
import os
import zlib

import numpy as np
import pytest

from nanoraw_b200 import _lib, _libio, fast5_io, synth
from nanoraw_b200 import nanoraw_helper as nh
from nanoraw_b200.frontend import genomeLoc, readInfo

EVENT_DTYPE = np.dtype([('norm_mean', '<f8'), ('norm_stdev', '<f8'), ('start', '<u4'), ('length', '<u4'),
                        ('base', 'S1')])


def test_library_exports_every_declared_symbol():
    lib = _libio.load()
    header = open(os.path.join(os.path.dirname(__file__), '..', 'include', 'nrq_io.h')).read()
    for name in _libio.SYMBOLS:
        assert name + '(' in header and hasattr(lib, name)


@pytest.mark.parametrize('seed', range(6))
def test_deflate_round_trips_through_zlib(seed):
    rng = np.random.default_rng(seed)
    cases = [b'', b'x', b'ab' * 31, b'a' * 64, bytes(range(256)) * 8, b'\x00' * 100_000,
             rng.integers(0, 256, 50_000, dtype=np.uint8).tobytes(),
             np.cumsum(rng.integers(2, 30, 20_000)).astype('<i8').tobytes()]
    fib = [1, 1]
    while len(fib) < 32:
        fib.append(fib[-1] + fib[-2])
    cases.append(b''.join(bytes([i]) * min(f, 150_000) for i, f in enumerate(fib)))  # forces the 15-bit limit
    for _ in range(40):
        k = int(rng.integers(1, 257))
        p = rng.dirichlet(np.ones(k) * float(rng.choice([0.05, 0.5, 5.0])))
        cases.append(rng.choice(k, int(rng.integers(0, 6000)), p=p).astype(np.uint8).tobytes())
    for data in cases:
        z = _libio.deflate(data)
        assert zlib.decompress(z) == data
        assert len(z) <= _libio.load().nrqio_deflate_bound(len(data))
        assert _libio.load().nrqio_adler32(np.frombuffer(data or b'\0', np.uint8).ctypes.data, len(data)) == \
            zlib.adler32(data)


def fake_results(reads, rng):
    """host CSR batch + plausible result arrays for it (what nrq_resquiggle_batch_host fills)"""
    from nanoraw_b200.batch import ReadBatch
    batch = ReadBatch.from_reads(reads)
    R, A = batch.n_reads, batch.total_align
    arrs = {f: np.ascontiguousarray(getattr(batch, f)) for f in ReadBatch.FIELDS}
    hb = _lib.NrqBatch()
    hb.n_reads = R
    for f, a in arrs.items():
        setattr(hb, f, a.ctypes.data)
    outs = {'status': np.zeros(R, np.int32), 'n_bases': np.zeros(R, np.int32), 'n_indels': np.zeros(R, np.int32),
            'scale_values': np.zeros(4 * R), 'new_segs': np.zeros(A + R, np.int32), 'norm_mean': rng.normal(size=A),
            'norm_stdev': np.abs(rng.normal(size=A)), 'ev_start': rng.integers(0, 1 << 20, A).astype(np.uint32),
            'ev_length': rng.integers(1, 60, A).astype(np.uint32),
            'ev_base': rng.choice(np.frombuffer(b'ACGT', np.uint8), A), 'region_samples': np.zeros(R, np.int32)}
    for r in range(R):
        g = np.frombuffer(reads[r].genome_align, np.uint8)
        outs['n_bases'][r] = int((g != ord('-')).sum())
    hr = _lib.NrqResults()
    for k, v in outs.items():
        setattr(hr, k, v.ctypes.data)
    return batch, arrs, hb, outs, hr


@pytest.mark.parametrize('threads', [1, 4])
def test_pack_corrected_matches_numpy(threads):
    rng = np.random.default_rng(11)
    reads = synth.make_reads(7, synth.SynthConfig(span_mean=300, span_sd=80, seed_base=81000),
                             genome=synth.make_genome(50_000, seed=2))
    batch, arrs, hb, outs, hr = fake_results(reads, rng)
    outs['status'][3] = _lib.ST_ZERO_LEN  # a failed read gets four empty chunks
    buf, off, ln = _libio.pack_corrected(hb, hr, batch.n_reads, threads=threads)
    for r, rd in enumerate(reads):
        parts = [buf[off[4 * r + j]:off[4 * r + j] + ln[4 * r + j]].tobytes() for j in range(4)]
        if r == 3:
            assert all(p == b'' for p in parts)
            continue
        o, n = int(batch.align_off[r]), int(outs['n_bases'][r])
        ev = np.frombuffer(zlib.decompress(parts[0]), dtype=EVENT_DTYPE)
        assert len(ev) == n
        for col, name in (('norm_mean', 'norm_mean'), ('norm_stdev', 'norm_stdev'), ('start', 'ev_start'),
                          ('length', 'ev_length')):
            assert np.array_equal(ev[col], outs[name][o:o + n])
        assert ev['base'].tobytes() == outs['ev_base'][o:o + n].tobytes()
        assert zlib.decompress(parts[1]) == rd.read_align and zlib.decompress(parts[2]) == rd.genome_align
        assert np.array_equal(np.frombuffer(zlib.decompress(parts[3]), '<i8'), rd.starts)
    assert _libio.load().nrqio_pack_corrected(C.byref(hb), C.byref(hr), 0, batch.n_reads, 1, buf.ctypes.data, 10,
                                              off.ctypes.data, ln.ctypes.data) == -1  # output too small
    # without the start column (not copied back from the device): rebuilt as the running sum of length from starts[0]
    hr.ev_start = None
    buf, off, ln = _libio.pack_corrected(hb, hr, batch.n_reads, threads=threads)
    for r, rd in enumerate(reads):
        if r == 3:
            continue
        o, n = int(batch.align_off[r]), int(outs['n_bases'][r])
        ev = np.frombuffer(zlib.decompress(buf[off[4 * r]:off[4 * r] + ln[4 * r]].tobytes()), dtype=EVENT_DTYPE)
        want = int(rd.starts[0]) + np.concatenate([[0], np.cumsum(outs['ev_length'][o:o + n][:-1], dtype=np.uint64)])
        assert np.array_equal(ev['start'], want.astype(np.uint32)) and np.array_equal(ev['length'], outs['ev_length'][o:o + n])


def _tree(node):
    out = {'attrs': {k: (np.asarray(v).tolist()) for k, v in node.attrs.items()}}
    if hasattr(node, 'keys'):
        out['children'] = {k: _tree(node[k]) for k in sorted(node.keys())}
    else:
        out['data'] = fast5_io.read_dataset(node)
        out['dtype'] = str(np.asarray(out['data']).dtype)
    return out


def _same_tree(a, b):
    assert a['attrs'] == b['attrs'] and ('children' in a) == ('children' in b)
    if 'children' in a:
        assert list(a['children']) == list(b['children'])
        for k in a['children']:
            _same_tree(a['children'][k], b['children'][k])
    else:
        assert a['dtype'] == b['dtype'] and a['data'].shape == b['data'].shape
        assert a['data'].tobytes() == b['data'].tobytes()


@pytest.mark.parametrize('backend', ['native', 'mem'])
def test_packed_write_back_equals_the_array_write_back(tmp_path, backend, monkeypatch):
    """write_packed_fast5_group (gzip chunks made by libnrqio) leaves the same RawGenomeCorrected tree as
    write_new_fast5_group (resquiggle.py:80-135) does from arrays"""
    from nanoraw_b200 import resquiggle as rsq
    monkeypatch.setattr(fast5_io, '_BACKEND', [backend])
    rng = np.random.default_rng(5)
    read = synth.make_reads(1, synth.SynthConfig(span_mean=400, span_sd=10, seed_base=82000),
                            genome=synth.make_genome(30_000, seed=2))[0]
    g = np.frombuffer(read.genome_align, np.uint8)
    n = int((g != ord('-')).sum())
    lengths = rng.integers(2, 30, n)
    ev = np.empty(n, dtype=EVENT_DTYPE)
    ev['norm_mean'], ev['norm_stdev'] = rng.normal(size=n), np.abs(rng.normal(size=n))
    ev['length'] = lengths
    ev['start'] = np.concatenate([[0], np.cumsum(lengths)[:-1]])
    ev['base'] = g[g != ord('-')].view('S1')
    new_segs = np.concatenate([ev['start'], [ev['start'][-1] + ev['length'][-1]]]).astype(np.int64)
    sv = nh.scaleValues(512.5, 41.0, -5.0, 5.0)
    loc = genomeLoc(1234, '-', 'chr1')
    info = readInfo('r', 'BaseCalled_template', 3, 4, 5, 6, 700, 80)
    names = []
    for k in range(2):
        fn = ('mem://packed_%d' % k) if backend == 'mem' else str(tmp_path / ('f%d.fast5' % k))
        synth.write_basecalled_fast5_job((fn, read))
        names.append(fn)
    rsq.write_new_fast5_group(names[0], loc, info, 77, new_segs, ev['base'].tobytes().decode(),
                              (read.read_align, read.genome_align), read.starts, None, sv,
                              'RawGenomeCorrected_000', 'BaseCalled_template', 'median', 5, True, event_data=ev)
    packed = (n, len(read.read_align), len(read.starts), _libio.deflate(ev.tobytes()),
              _libio.deflate(read.read_align), _libio.deflate(read.genome_align),
              _libio.deflate(np.asarray(read.starts, '<i8').tobytes()))
    rsq.write_packed_fast5_group(names[1], loc, info, 77, packed, sv, 'RawGenomeCorrected_000',
                                 'BaseCalled_template', 'median', 5)
    trees = []
    for fn in names:
        f = fast5_io.open_fast5(fn, 'r')
        trees.append(_tree(f['/Analyses/RawGenomeCorrected_000']))
        f.close()
    _same_tree(trees[0], trees[1])
    # h5py refuses a None attribute: the reference's reads with --outlier-threshold <= 0 fail in the write (:94-97)
    with pytest.raises(NotImplementedError):
        fn = 'mem://packed_none' if backend == 'mem' else str(tmp_path / 'none.fast5')
        synth.write_basecalled_fast5_job((fn, read))
        rsq.write_packed_fast5_group(fn, loc, info, 77, packed, sv, 'RawGenomeCorrected_000',
                                     'BaseCalled_template', 'median', None)


def test_h5py_reads_what_the_native_writer_and_the_packer_produce(tmp_path, monkeypatch):
    """cross-check against the real HDF5 library wherever it is installed (it is not in the build image): files
    written by hdf5_min.py, with datasets packed by libnrqio, opened with h5py"""
    h5py = pytest.importorskip('h5py')
    from nanoraw_b200 import resquiggle as rsq
    monkeypatch.setattr(fast5_io, '_BACKEND', ['native'])
    rng = np.random.default_rng(6)
    read = synth.make_reads(1, synth.SynthConfig(span_mean=300, span_sd=10, seed_base=83000),
                            genome=synth.make_genome(30_000, seed=2))[0]
    fn = str(tmp_path / 'x.fast5')
    synth.write_basecalled_fast5_job((fn, read))
    g = np.frombuffer(read.genome_align, np.uint8)
    n = int((g != ord('-')).sum())
    ev = np.zeros(n, dtype=EVENT_DTYPE)
    ev['norm_mean'] = rng.normal(size=n)
    ev['base'] = g[g != ord('-')].view('S1')
    packed = (n, len(read.read_align), len(read.starts), _libio.deflate(ev.tobytes()),
              _libio.deflate(read.read_align), _libio.deflate(read.genome_align),
              _libio.deflate(np.asarray(read.starts, '<i8').tobytes()))
    rsq.write_packed_fast5_group(fn, genomeLoc(5, '+', 'c'), readInfo('r', 'BaseCalled_template', 0, 0, 1, 2, 3, 4),
                                 9, packed, nh.scaleValues(1.0, 2.0, -5.0, 5.0), 'RawGenomeCorrected_000',
                                 'BaseCalled_template', 'median', 5)
    with h5py.File(fn, 'r') as f:
        assert np.array_equal(f['/Raw/Reads/Read_1/Signal'][()], read.raw)
        grp = f['/Analyses/RawGenomeCorrected_000/BaseCalled_template']
        assert np.array_equal(grp['Events'][()], ev) and grp['Events'].attrs['read_start_rel_to_raw'] == 9
        assert grp['Alignment/read_alignment'][()].tobytes() == read.read_align
        assert np.array_equal(grp['Alignment/read_segments'][()], read.starts)
        assert grp.attrs['shift'] == 1.0 and grp['Alignment'].attrs['mapped_chrom'] in ('c', b'c')


# ----------------------------------------------------------------------------------------------- nrqio_inflate
def _zlib_streams(data):
    """the same bytes as every kind of deflate stream zlib can write, plus this library's own encoder"""
    for level in (0, 1, 4, 6, 9):
        yield zlib.compress(data, level)
    for strategy in (zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED):
        c = zlib.compressobj(6, zlib.DEFLATED, 15, 8, strategy)
        yield c.compress(data) + c.flush()
    c = zlib.compressobj(6, zlib.DEFLATED, 9, 1)      # 512-byte window, smallest hash
    yield c.compress(data) + c.flush()
    c, out = zlib.compressobj(5), b''
    for k in range(0, len(data), 1000):               # many blocks, with empty stored blocks between them
        out += c.compress(data[k:k + 1000]) + c.flush(zlib.Z_SYNC_FLUSH if (k // 1000) % 2 else zlib.Z_FULL_FLUSH)
    yield out + c.flush()
    yield _libio.deflate(data)


def test_inflate_equals_zlib_on_every_kind_of_stream():
    rng = np.random.default_rng(0)
    cases = [b'', b'a', b'abc' * 1000, bytes(range(256)) * 50, b'\x00' * 70000,
             rng.integers(0, 256, 100000, dtype=np.uint8).tobytes(),
             np.cumsum(rng.integers(-30, 31, 90000)).astype('<i2').tobytes(), (b'ACGT' * 3 + b'TTAG') * 5000,
             rng.integers(0, 4, 50000, dtype=np.uint8).tobytes()]
    cases += [r.raw.tobytes() for r in synth.make_reads(2, synth.config_short())]
    for _ in range(40):
        n, k = int(rng.integers(0, 3000)), int(rng.integers(1, 257))
        cases.append(rng.choice(k, n, p=rng.dirichlet(np.ones(k) * 0.3)).astype(np.uint8).tobytes())
    for data in cases:
        for z in _zlib_streams(data):
            assert zlib.decompress(z) == data
            assert _libio.inflate(z, len(data)) == data
            assert _libio.inflate(z, len(data) + 300) == data
            if data:
                assert _libio.inflate(z, len(data) - 1) is None      # more bytes than the caller has room for
            for cut in (1, 2, 5, len(z) // 2, len(z) - 5, len(z) - 1):
                if 0 < cut < len(z):
                    assert _libio.inflate(z[:cut], len(data) + 10) is None
            if len(z) > 12:                                           # one flipped bit: refused (or, with a stored
                zz = bytearray(z)                                     # block's padding bits, harmless)
                zz[int(rng.integers(2, len(z) - 4))] ^= 1 << int(rng.integers(0, 8))
                assert _libio.inflate(bytes(zz), len(data) + 10) in (None, data)


def test_inflate_refuses_garbage_without_touching_memory_outside():
    rng = np.random.default_rng(1)
    lib = _libio.load()
    for k in range(3000):
        n = int(rng.integers(0, 400))
        junk = rng.integers(0, 256, n, dtype=np.uint8)
        if n >= 2 and k % 2:
            junk[0], junk[1] = 0x78, 0x9c                             # a valid zlib header in front of the noise
        cap = int(rng.integers(0, 600))
        dst = np.full(cap + 64, 0xAB, dtype=np.uint8)
        src = np.concatenate([junk, np.zeros(1, np.uint8)])           # (so that an empty input has an address)
        got = int(lib.nrqio_inflate(src.ctypes.data, n, dst[32:].ctypes.data, cap))
        assert got == -1 or 0 <= got <= cap
        assert (dst[:32] == 0xAB).all() and (dst[32 + cap:] == 0xAB).all()


def test_adler32_equals_zlib():
    """the vector path (32 bytes per step, runs of 173 steps) against zlib: lengths around every boundary of the
    algorithm, the contents that overflow first, unaligned starts"""
    lib = _libio.load()
    rng = np.random.default_rng(2)
    for n in [0, 1, 2, 31, 32, 33, 63, 64, 65, 100, 5535, 5536, 5537, 5552, 5553, 11072, 11073, 100000, 242000, (1 << 21) + 7]:
        for kind in range(3):
            data = rng.integers(0, 256, n, dtype=np.uint8) if kind == 0 else np.full(n, 255 if kind == 1 else 0, np.uint8)
            buf = np.concatenate([data, np.zeros(1, np.uint8)])
            assert lib.nrqio_adler32(buf.ctypes.data, n) == zlib.adler32(data.tobytes()), (n, kind)
    for _ in range(200):
        n, off = int(rng.integers(0, 20000)), int(rng.integers(0, 40))
        base = rng.integers(0, 256, n + off + 1, dtype=np.uint8)
        assert lib.nrqio_adler32(base[off:].ctypes.data, n) == zlib.adler32(base[off:off + n].tobytes())
