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


def test_params_struct_layout(tmp_path):
    """the ctypes mirrors against the C compiler's view of include/nrq.h: size and every field offset"""
    import re
    import subprocess
    assert C.sizeof(_lib.NrqParams) == 56 and _lib.NrqParams.timeout_ns.offset == 32
    assert C.sizeof(_lib.NrqBatch) == 96 and C.sizeof(_lib.NrqResults) == 88
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    structs = {'nrq_params': _lib.NrqParams, 'nrq_batch': _lib.NrqBatch, 'nrq_results': _lib.NrqResults}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "nrq.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['return 0; }']
    src = tmp_path / 'probe.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'probe'
    subprocess.run(['gcc', '-I', os.path.join(root, 'include'), '-o', str(exe), str(src)], check=True)
    out = dict(ln.split() for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True)
               .stdout.splitlines())
    for cname, cls in structs.items():
        assert int(out[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(out['%s.%s' % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)


def test_engine_refuses_without_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from nanoraw_b200.engine import ResquiggleEngine
    with pytest.raises(_lib.NrqError):
        ResquiggleEngine()
