"""ctypes binding of libnrq.so (include/nrq.h).  There is no CPU fallback: if the shared
library is missing or cannot be loaded, importing the compute path fails loudly."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('NRQ_LIB_PATH') or os.path.join(_HERE, 'libnrq.so')

# status codes (include/nrq.h)
ST_OK, ST_CPTS_LIMIT, ST_TIMEOUT, ST_ZERO_LEN, ST_NEG_START, ST_PAST_END, ST_SEQ_MISMATCH, \
    ST_FPE_DIVIDE, ST_FPE_INVALID, ST_UNSATISFIABLE, ST_REGION_TOO_LARGE, ST_BAD_ALIGNMENT, \
    ST_BAD_INPUT, ST_WORKSPACE = range(14)
NORM_MEDIAN, NORM_GIVEN = 0, 1


class NrqParams(C.Structure):
    _fields_ = [
        ('norm_type', C.c_int32), ('use_outlier', C.c_int32), ('outlier_thresh', C.c_double),
        ('given_limits', C.c_int32), ('compute_sd', C.c_int32), ('num_cpts_limit', C.c_int32),
        ('min_seg_len', C.c_int32), ('timeout_ns', C.c_int64), ('big_region_cap', C.c_int32),
        ('long_read_parts', C.c_int32), ('debug_flags', C.c_int32)]


class NrqBatch(C.Structure):
    _fields_ = [
        ('n_reads', C.c_int32), ('raw', C.c_void_p), ('raw_off', C.c_void_p),
        ('read_start', C.c_void_p), ('starts', C.c_void_p), ('starts_off', C.c_void_p),
        ('read_align', C.c_void_p), ('genome_align', C.c_void_p), ('align_off', C.c_void_p),
        ('scale_values_in', C.c_void_p), ('norm_signal', C.c_void_p), ('norm_off', C.c_void_p)]


class NrqResults(C.Structure):
    _fields_ = [
        ('status', C.c_void_p), ('n_bases', C.c_void_p), ('n_indels', C.c_void_p),
        ('scale_values', C.c_void_p), ('new_segs', C.c_void_p), ('norm_mean', C.c_void_p),
        ('norm_stdev', C.c_void_p), ('ev_start', C.c_void_p), ('ev_length', C.c_void_p),
        ('ev_base', C.c_void_p), ('region_samples', C.c_void_p)]


# every symbol include/nrq.h declares: (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    'nrq_version': (C.c_int, []),
    'nrq_last_error': (C.c_char_p, []),
    'nrq_status_message': (C.c_char_p, [C.c_int32]),
    'nrq_workspace_bytes': (C.c_int64, [C.c_int32, C.c_int64, C.POINTER(NrqParams)]),
    'nrq_find_indels': (C.c_int, [C.POINTER(NrqBatch), _P, _P, _P, _P, _P, C.c_int64, _P, _P]),
    'nrq_norm_stats': (C.c_int, [C.POINTER(NrqBatch), C.POINTER(NrqParams), _P, _P, _P]),
    'nrq_normalize_signal': (C.c_int, [C.POINTER(NrqBatch), _P, _P, _P, _P, _P]),
    'nrq_resegment': (C.c_int, [C.POINTER(NrqBatch), C.POINTER(NrqParams), _P, _P, _P, _P, _P, _P, C.c_int64,
                                _P, C.c_int64, _P, _P, _P, _P, _P, _P]),
    'nrq_running_diffs': (C.c_int, [_P, C.c_int32, C.c_int32, _P, _P, _P]),
    'nrq_genome_aggregate': (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P,
                                       _P, _P, _P]),
    'nrq_genome_event_lists': (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P]),
    'nrq_event_table': (C.c_int, [C.POINTER(NrqBatch), C.POINTER(NrqParams), _P, _P, _P, _P, _P, _P,
                                  _P, _P, _P, _P]),
    'nrq_resquiggle_batch': (C.c_int, [C.POINTER(NrqBatch), C.POINTER(NrqParams),
                                       C.POINTER(NrqResults), _P, C.c_int64, _P]),
    'nrq_resquiggle_batch_profile': (C.c_int, [C.POINTER(NrqBatch), C.POINTER(NrqParams),
                                               C.POINTER(NrqResults), _P, C.c_int64, _P,
                                               C.POINTER(C.c_float)]),
    'nrq_create': (_P, [C.c_int32]),
    'nrq_destroy': (None, [_P]),
    'nrq_resquiggle_batch_host': (C.c_int, [_P, C.POINTER(NrqBatch), C.POINTER(NrqParams),
                                            C.POINTER(NrqResults)]),
    'nrq_resquiggle_batch_host_submit': (C.c_int64, [_P, C.POINTER(NrqBatch), C.POINTER(NrqParams),
                                                     C.POINTER(NrqResults)]),
    'nrq_resquiggle_batch_host_wait': (C.c_int, [_P, C.c_int64]),
    'nrq_launch_count': (C.c_int64, []),
}

_lib = None


class NrqError(RuntimeError):
    pass


def load():
    """Load libnrq.so (built by nanoraw_b200.csrc.build); raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NrqError(
            'libnrq.so not found at %s: build it with `python -m nanoraw_b200.csrc.build` '
            '(the CUDA path has no CPU fallback)' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        raise NrqError('%s failed (%d): %s' % (what, rc, load().nrq_last_error().decode()))


def status_message(code):
    return load().nrq_status_message(int(code)).decode()
