"""ctypes binding of libnrqio.so (include/nrq_io.h): the host-side byte work of the FAST5 write-back (deflate
without match search, packing of a result batch into gzip chunks).  Plain C++ (no CUDA): the I/O processes may
load it too.  Like libnrq it has no Python fallback on the product path: packing fails loudly if it is missing."""
import ctypes as C
import os

import numpy as np

from . import _lib

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('NRQIO_LIB_PATH') or os.path.join(_HERE, 'libnrqio.so')

_P = C.c_void_p
SYMBOLS = {
    'nrqio_deflate_bound': (C.c_int64, [C.c_int64]),
    'nrqio_deflate': (C.c_int64, [_P, C.c_int64, _P, C.c_int64]),
    'nrqio_adler32': (C.c_uint32, [_P, C.c_int64]),
    'nrqio_inflate': (C.c_int64, [_P, C.c_int64, _P, C.c_int64]),
    'nrqio_pack_corrected_bound': (C.c_int64, [C.POINTER(_lib.NrqBatch), C.POINTER(_lib.NrqResults), C.c_int32,
                                               C.c_int32]),
    'nrqio_pack_corrected': (C.c_int, [C.POINTER(_lib.NrqBatch), C.POINTER(_lib.NrqResults), C.c_int32, C.c_int32,
                                       C.c_int32, _P, C.c_int64, _P, _P]),
}

SYMBOLS.update({
    'nrqio_fe_message': (C.c_char_p, [C.c_int32]),
    'nrqio_events_to_starts': (C.c_int, [_P, C.c_int32, _P, C.c_int32, _P, C.c_int32, _P, C.c_int32, C.c_int64,
                                         C.c_int32, C.c_int64, C.c_uint64, _P, _P, _P, _P]),
    'nrqio_sam_to_rows': (C.c_int, [C.c_char_p, C.c_int32, C.c_int64, C.c_char_p, C.c_int64, C.c_char_p, C.c_int64,
                                    _P, _P, C.c_int64, _P, _P, _P]),
    'nrqio_clip_starts': (C.c_int, [_P, _P, _P, C.c_int64, C.c_int64]),
    'nrqio_frontend_batch': (C.c_int, [_P, C.c_int32, C.c_int32, _P, _P, _P, C.c_int64, _P, _P, _P, C.c_int64, _P,
                                       _P, _P]),
})



class F5CorrectedBatch(C.Structure):
    """nrqio_f5_corrected_batch"""
    _fields_ = [('n', C.c_int32), ('outlier_is_int', C.c_int32), ('corrected_group', C.c_char_p),
                ('subgroup', C.c_char_p), ('norm_type', C.c_char_p), ('outlier_threshold', C.c_double),
                ('strtab', _P), ('path_off', _P), ('chrom_off', _P), ('strand_off', _P), ('scale_values', _P),
                ('ints', _P), ('n_bases', _P), ('n_align', _P), ('n_segs', _P), ('chunks', _P), ('chunk_off', _P),
                ('chunk_len', _P), ('skip', _P)]


SYMBOLS.update({
    'nrqio_f5_open_raw_batch': (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.POINTER(_P), _P, _P, _P]),
    'nrqio_f5_read_raw_batch': (C.c_int, [_P, _P, _P, C.c_int32, _P]),
    'nrqio_f5_close_batch': (None, [_P]),
    'nrqio_f5_write_corrected_batch': (C.c_int, [C.POINTER(F5CorrectedBatch), C.c_int32, _P]),
})

F5_OK, F5_IO, F5_FORMAT, F5_UNSUPPORTED, F5_MISSING, F5_EXISTS = range(6)

FE_OK, FE_EVENTS_BEFORE_RAW, FE_TOO_FEW_SEGMENTS, FE_ZERO_LENGTH, FE_ALL_STAYS, FE_BAD_CIGAR, \
    FE_SEQ_CIGAR_DISAGREE, FE_CAPACITY, FE_BAD_INPUT = range(9)


class FeRead(C.Structure):
    """nrqio_fe_read"""
    _fields_ = [('cigar', C.c_char_p), ('flag', C.c_int32), ('pos', C.c_int64), ('seq', C.c_char_p),
                ('seq_len', C.c_int64), ('contig', C.c_char_p), ('contig_len', C.c_int64), ('starts', _P),
                ('n_starts', C.c_int64), ('read_start', C.c_int64)]


_io = None


def load():
    global _io
    if _io is not None:
        return _io
    if not os.path.exists(LIB_PATH):
        raise _lib.NrqError('libnrqio.so not found at %s: build it with `python -m nanoraw_b200.csrc.build`'
                            % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _io = lib
    return lib


def deflate(data):
    """zlib stream (what an HDF5 gzip chunk holds) of `data` (bytes-like), by nrqio_deflate"""
    lib = load()
    src = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, np.uint8)
    cap = int(lib.nrqio_deflate_bound(len(data)))
    dst = np.empty(cap, dtype=np.uint8)
    n = int(lib.nrqio_deflate(src.ctypes.data, len(data), dst.ctypes.data, cap))
    if n < 0:
        raise _lib.NrqError('nrqio_deflate failed')
    return dst[:n].tobytes()


def inflate(data, size):
    """bytes of the zlib stream `data`, which must hold at most `size` of them (nrqio_inflate); None if refused"""
    lib = load()
    src = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, np.uint8)
    dst = np.empty(max(int(size), 1), dtype=np.uint8)
    n = int(lib.nrqio_inflate(src.ctypes.data, len(data), dst.ctypes.data, int(size)))
    return None if n < 0 else dst[:n].tobytes()


def pack_threads():
    env = os.environ.get('NRQ_PACK_THREADS')
    if env is not None:
        return max(1, int(env))
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    return max(1, min(16, cores))


def pack_corrected(c_batch, c_results, n_reads, threads=None):
    """the four gzip chunks (Events, read_alignment, genome_alignment, read_segments) of every read of a host
    result batch -> (buffer uint8[], chunk_off int64[4 R], chunk_len int64[4 R])"""
    lib = load()
    bound = int(lib.nrqio_pack_corrected_bound(C.byref(c_batch), C.byref(c_results), 0, n_reads))
    if bound < 0:
        raise _lib.NrqError('nrqio_pack_corrected_bound failed')
    out = np.empty(max(bound, 1), dtype=np.uint8)
    off = np.zeros(4 * max(n_reads, 1), dtype=np.int64)
    ln = np.zeros(4 * max(n_reads, 1), dtype=np.int64)
    rc = lib.nrqio_pack_corrected(C.byref(c_batch), C.byref(c_results), 0, n_reads,
                                  pack_threads() if threads is None else int(threads),
                                  out.ctypes.data, out.size, off.ctypes.data, ln.ctypes.data)
    if rc != 0:
        raise _lib.NrqError('nrqio_pack_corrected failed')
    return out, off, ln


def fe_message(status):
    return load().nrqio_fe_message(int(status)).decode()


def events_to_starts(events, albacore_after_1_0, sampling_rate, raw_start_time):
    """nrqio_events_to_starts on an Events table -> (status, starts int64[], bases bytes, read_start)"""
    lib = load()
    n = len(events)
    cols = {}
    for name in ('start', 'length'):
        c = np.ascontiguousarray(events[name])
        if c.dtype not in (np.float32, np.float64):
            c = c.astype(np.float64)
        cols[name] = c
    states = np.ascontiguousarray(events['model_state'])
    if states.dtype.kind != 'S':
        states = states.astype('S')
    move = np.ascontiguousarray(events['move'])
    if move.dtype.kind not in 'iu' or move.dtype.kind == 'u':
        move = move.astype(np.int64)
    starts = np.empty(n + 1, dtype=np.int64)
    bases = np.empty(max(n, 1), dtype=np.uint8)
    n_starts = C.c_int64(0)
    read_start = C.c_int64(0)
    st = lib.nrqio_events_to_starts(
        cols['start'].ctypes.data, int(cols['start'].dtype == np.float32),
        cols['length'].ctypes.data, int(cols['length'].dtype == np.float32),
        states.ctypes.data, states.dtype.itemsize, move.ctypes.data, move.dtype.itemsize, n,
        int(bool(albacore_after_1_0)), int(sampling_rate), int(raw_start_time),
        starts.ctypes.data, bases.ctypes.data, C.byref(n_starts), C.byref(read_start))
    k = n_starts.value
    return st, starts[:k].copy(), bases[:max(k - 1, 0)].tobytes(), read_start.value


def sam_to_rows(cigar, flag, pos, seq, contig):
    """nrqio_sam_to_rows -> (status, read_row bytes, genome_row bytes, (clip_start, clip_end), tallies[4])"""
    lib = load()
    cigar_b = cigar.encode() if isinstance(cigar, str) else cigar
    seq_b = seq.encode() if isinstance(seq, str) else seq
    contig_b = contig.encode() if isinstance(contig, str) else contig
    cap = len(seq_b) + sum(int(x) for x in __import__('re').findall(rb'(\d+)[MIDNSHP=X]', cigar_b)) + 16
    rr = np.empty(cap, dtype=np.uint8)
    gr = np.empty(cap, dtype=np.uint8)
    n_cols = C.c_int64(0)
    clip = np.zeros(2, dtype=np.int64)
    tallies = np.zeros(4, dtype=np.int64)
    st = lib.nrqio_sam_to_rows(cigar_b, int(flag), int(pos), seq_b, len(seq_b), contig_b, len(contig_b),
                               rr.ctypes.data, gr.ctypes.data, cap, C.byref(n_cols), clip.ctypes.data,
                               tallies.ctypes.data)
    k = n_cols.value
    return st, rr[:k].tobytes(), gr[:k].tobytes(), (int(clip[0]), int(clip[1])), tallies


# ---------------------------------------------------------------------------------------------- FAST5 files
def f5_threads():
    """threads of the batch file functions (NRQ_F5_THREADS; default: the host cores, at most 32)"""
    env = os.environ.get('NRQ_F5_THREADS')
    if env is not None:
        return max(1, int(env))
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    return max(1, min(32, cores))


def _strtab(strings):
    """list of str / bytes -> (uint8 array of NUL-terminated strings, int64 offsets)"""
    raw = [x if isinstance(x, bytes) else os.fsencode(x) if isinstance(x, os.PathLike) else str(x).encode('utf-8')
           for x in strings]
    off = np.zeros(len(raw), dtype=np.int64)
    if raw:
        np.cumsum([len(x) + 1 for x in raw[:-1]], out=off[1:])
    tab = np.frombuffer(b'\x00'.join(raw) + b'\x00', dtype=np.uint8)
    return tab, off


class F5RawBatch(object):
    """nrqio_f5_open_raw_batch over a list of file names: `n_samples`, `channel` (offset, range, digitisation,
    sampling_rate per file) and `status` are known at once; `read_into` inflates the signals into a caller's
    int16 array (pinned staging); `close` drops the file images."""

    def __init__(self, paths, threads=None):
        lib = load()
        n = len(paths)
        self.n = n
        self.threads = f5_threads() if threads is None else int(threads)
        self.n_samples = np.zeros(max(n, 1), dtype=np.int64)[:n]
        self.channel = np.full((max(n, 1), 4), np.nan, dtype=np.float64)[:n]
        self.status = np.zeros(max(n, 1), dtype=np.int32)[:n]
        self._h = _P()
        tab, off = _strtab(paths)
        rc = lib.nrqio_f5_open_raw_batch(tab.ctypes.data, off.ctypes.data, n, self.threads, C.byref(self._h),
                                         self.n_samples.ctypes.data, self.channel.ctypes.data,
                                         self.status.ctypes.data)
        if rc != 0:
            raise _lib.NrqError('nrqio_f5_open_raw_batch failed')

    def read_into(self, dst, dst_off):
        """signals -> dst[dst_off[i] : dst_off[i] + n_samples[i]] (dst_off[i] < 0: skipped); -> status int32[n]"""
        if dst.dtype != np.int16 or not dst.flags['C_CONTIGUOUS']:
            raise ValueError('dst must be a contiguous int16 array')
        dst_off = np.ascontiguousarray(dst_off, dtype=np.int64)
        ok = (self.status == F5_OK) & (dst_off >= 0)
        if ok.any() and int((dst_off[ok] + self.n_samples[ok]).max()) > dst.size:
            raise ValueError('dst is too small')
        status = np.zeros(max(self.n, 1), dtype=np.int32)[:self.n]
        rc = load().nrqio_f5_read_raw_batch(self._h, dst.ctypes.data, dst_off.ctypes.data, self.threads,
                                            status.ctypes.data)
        if rc != 0:
            raise _lib.NrqError('nrqio_f5_read_raw_batch failed')
        return status

    def close(self):
        if self._h:
            load().nrqio_f5_close_batch(self._h)
            self._h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 -- interpreter shutdown
            pass


def f5_write_corrected(paths, corrected_group, subgroup, norm_type, outlier_thresh, chroms, strands, scale_values,
                       ints, n_bases, n_align, n_segs, chunks, chunk_off, chunk_len, skip=None, threads=None):
    """nrqio_f5_write_corrected_batch: RawGenomeCorrected groups appended to their files -> status int32[n].
    scale_values float64[n, 4]; ints int64[n, 8] (mapped_start, clipped_bases_start, clipped_bases_end,
    num_insertions, num_deletions, num_matches, num_mismatches, read_start_rel_to_raw); chunks / chunk_off /
    chunk_len as pack_corrected returns them."""
    lib = load()
    n = len(paths)
    tab, offs = _strtab(list(paths) + list(chroms) + list(strands))
    keep = [tab, offs]

    def arr(a, dt, count):
        a = np.ascontiguousarray(a, dtype=dt).reshape(-1)
        if a.size < count:
            raise ValueError('array too short')
        keep.append(a)
        return a.ctypes.data

    w = F5CorrectedBatch()
    w.n = n
    w.outlier_is_int = int(isinstance(outlier_thresh, (int, np.integer)) and not isinstance(outlier_thresh, bool))
    w.corrected_group = corrected_group.encode()
    w.subgroup = subgroup.encode()
    w.norm_type = norm_type.encode()
    w.outlier_threshold = float(outlier_thresh)
    w.strtab = tab.ctypes.data
    w.path_off = arr(offs[:n], np.int64, n)
    w.chrom_off = arr(offs[n:2 * n], np.int64, n)
    w.strand_off = arr(offs[2 * n:3 * n], np.int64, n)
    w.scale_values = arr(scale_values, np.float64, 4 * n)
    w.ints = arr(ints, np.int64, 8 * n)
    w.n_bases = arr(n_bases, np.int32, n)
    w.n_align = arr(n_align, np.int64, n)
    w.n_segs = arr(n_segs, np.int64, n)
    w.chunks = chunks.ctypes.data
    w.chunk_off = arr(chunk_off, np.int64, 4 * n)
    w.chunk_len = arr(chunk_len, np.int64, 4 * n)
    w.skip = arr(skip, np.int32, n) if skip is not None else None
    status = np.zeros(max(n, 1), dtype=np.int32)[:n]
    rc = lib.nrqio_f5_write_corrected_batch(C.byref(w), f5_threads() if threads is None else int(threads),
                                            status.ctypes.data)
    if rc != 0:
        raise _lib.NrqError('nrqio_f5_write_corrected_batch failed')
    return status
