"""The C-ABI library loads and exports every symbol include/b200fe.h declares (no GPU needed)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "b200fe.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200fe_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import benchmarks_b200 as b
    syms = _declared_symbols()
    assert len(syms) >= 7
    lib = ctypes.CDLL(b.LIB_PATH)
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in b200fe.h but not exported: {missing}"
    assert b.lib.b200fe_version() == 100


def test_python_binding_covers_header():
    from benchmarks_b200 import _lib
    assert sorted(_lib._SIGNATURES) == _declared_symbols()


def test_argument_errors_are_reported_not_thrown():
    import benchmarks_b200 as b
    # degree outside 1..8 and wrong nq: B200FE_ERR_UNSUPPORTED (2) before any CUDA call
    assert b.lib.b200fe_bk5_apply(0, 1, None, None, None, None, None) == 2
    assert b.lib.b200fe_bk5_apply(9, 1, None, None, None, None, None) == 2
    assert b.lib.b200fe_bk3_apply(3, 4, 1, None, None, None, None, None, None) == 2
    assert b"nq=4" in b.lib.b200fe_last_error()
    # null pointers: B200FE_ERR_INVALID_ARG (1)
    assert b.lib.b200fe_bk1_apply(2, 4, 1, None, None, None, None, None) == 1
