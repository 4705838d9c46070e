"""The C-ABI library loads and exports every symbol include/psbax_b200.h declares; argument
validation that needs no GPU returns the documented error codes."""
import ctypes as C
import os
import re

from helpers import ROOT

from psbax_b200 import _lib


def _declared():
    text = open(os.path.join(ROOT, "include", "psbax_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(psbax_\w+)\s*\(", text)))


def test_header_symbols_are_exported():
    lib = _lib.load()
    names = _declared()
    assert names, "no declarations parsed"
    assert sorted(_lib.EXPORTS) == names
    for n in names:
        assert hasattr(lib, n), n


def test_version_and_struct_size():
    lib = _lib.load()
    assert lib.psbax_version() == 2
    # 8 int32 + 6 pointers + 4 floats
    assert C.sizeof(_lib.PathsStruct) == 8 * 4 + 6 * 8 + 4 * 4
    assert C.sizeof(_lib.MeshStruct) == 2 * 4 + 4 * 8 + 4 * 8
    # the product build keeps no trace state (include/psbax_b200.h: no global mutable state)
    assert lib.psbax_debug_trace_buffer(None) == -2


def test_invalid_arguments_report_errors():
    lib = _lib.load()
    n = C.c_size_t(0)
    assert lib.psbax_pack_bytes(None, C.byref(n)) == -1
    assert "NULL" in _lib.last_error()
    s = _lib.PathsStruct(d=2, L=8, G=3, S=4, n=0, kernel=3, engine=0)
    s.omega = s.bias = s.W = 1  # non-NULL dummies; nothing is dereferenced during validation
    assert lib.psbax_pack_bytes(C.byref(s), C.byref(n)) == -1
    assert "G=3" in _lib.last_error()
    s.G = 4
    s.kernel = 9
    assert lib.psbax_pack_bytes(C.byref(s), C.byref(n)) == -1
    s.kernel = 3
    assert lib.psbax_pack_bytes(C.byref(s), C.byref(n)) == 0 and n.value > 0
    s.engine = 2  # tensor engine needs a shared basis
    assert lib.psbax_pack_bytes(C.byref(s), C.byref(n)) == -2
    s.engine = 3
    assert lib.psbax_pack_bytes(C.byref(s), C.byref(n)) == -2
