"""Per-read status codes -> the exceptions the reference raises.

The device path never throws per read; it fills status[r].  The host mirror of
resquiggle_read turns a status back into the exception type and message the reference
would have produced, so failed_reads summaries stay byte-identical.
"""
from . import _lib

STATUS_EXCEPTIONS = {
    _lib.ST_CPTS_LIMIT: RuntimeError,          # resquiggle.py:235
    _lib.ST_TIMEOUT: RuntimeError,             # resquiggle.py:289
    _lib.ST_ZERO_LEN: NotImplementedError,     # resquiggle.py:374
    _lib.ST_NEG_START: NotImplementedError,    # resquiggle.py:377
    _lib.ST_PAST_END: NotImplementedError,     # resquiggle.py:380
    _lib.ST_SEQ_MISMATCH: ValueError,          # resquiggle.py:386
    _lib.ST_FPE_DIVIDE: FloatingPointError,    # np.seterr(all='raise'), resquiggle.py:8
    _lib.ST_FPE_INVALID: FloatingPointError,
    _lib.ST_UNSATISFIABLE: RuntimeError,       # the reference hangs here
    _lib.ST_REGION_TOO_LARGE: RuntimeError,
    _lib.ST_BAD_ALIGNMENT: RuntimeError,
    _lib.ST_BAD_INPUT: RuntimeError,
    _lib.ST_WORKSPACE: RuntimeError,
}


def exception_for(status):
    """None for NRQ_ST_OK, else an exception instance carrying the reference's message."""
    status = int(status)
    if status == _lib.ST_OK:
        return None
    return STATUS_EXCEPTIONS.get(status, RuntimeError)(_lib.status_message(status))


def error_string(status):
    """'ExceptionType: message' ('' when OK) -- the form tests/golden stores."""
    e = exception_for(status)
    return '' if e is None else '%s: %s' % (type(e).__name__, e)
