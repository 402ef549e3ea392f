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
