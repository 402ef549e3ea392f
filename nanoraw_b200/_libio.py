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
    return max(1, min(8, cores // 2))


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
