"""CPU-side checks of the C-ABI library: it loads and exports every symbol include/nrq.h
declares.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re

from nanoraw_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    from nanoraw_b200.csrc import build
    return build.build()


def test_library_builds_and_loads():
    path = _build()
    assert os.path.exists(path)
    lib = _lib.load()
    assert lib.nrq_version() == 1


def test_every_declared_symbol_is_exported():
    _build()
    header = open(os.path.join(ROOT, 'include', 'nrq.h')).read()
    declared = set(re.findall(r'\b(nrq_[a-z_]+)\s*\(', header))
    assert declared, 'no declarations found'
    lib = C.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))


def test_status_messages_match_reference_strings():
    _build()
    from oracle import resquiggle_oracle as orc
    assert _lib.status_message(_lib.ST_CPTS_LIMIT) == orc.MSG_CPTS_LIMIT
    assert _lib.status_message(_lib.ST_TIMEOUT) == orc.MSG_TIMEOUT
    assert _lib.status_message(_lib.ST_ZERO_LEN) == orc.MSG_ZERO_LEN
    assert _lib.status_message(_lib.ST_NEG_START) == orc.MSG_NEG_START
    assert _lib.status_message(_lib.ST_PAST_END) == orc.MSG_PAST_END
    assert _lib.status_message(_lib.ST_SEQ_MISMATCH) == orc.MSG_SEQ_MISMATCH
    assert _lib.status_message(_lib.ST_UNSATISFIABLE) == orc.MSG_UNSATISFIABLE
    assert _lib.status_message(_lib.ST_OK) == ''


def test_params_struct_layout():
    # must match struct nrq_params in include/nrq.h (natural alignment)
    assert C.sizeof(_lib.NrqParams) == 48
    assert _lib.NrqParams.outlier_thresh.offset == 8
    assert _lib.NrqParams.timeout_ns.offset == 32
    assert C.sizeof(_lib.NrqBatch) == 96
    assert C.sizeof(_lib.NrqResults) == 88


def test_engine_refuses_without_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from nanoraw_b200.engine import ResquiggleEngine
    with pytest.raises(_lib.NrqError):
        ResquiggleEngine()
