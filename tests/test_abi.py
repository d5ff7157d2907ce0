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


def test_no_cpu_fallback_without_a_device():
    """On a machine without a CUDA device every compute entry point returns B200FE_ERR_CUDA (3) -- it never computes on the
    host.  (Skipped where a GPU is present: there the same calls succeed and are covered by the -m gpu tests.)"""
    import numpy as np
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import ctypes as C
    import benchmarks_b200 as b
    p, nq, nelmt = 2, 4, 3
    nm = p + 1
    basis, dbasis = np.ones(nq * nm), np.ones(nq * nq)
    buf = np.zeros(nelmt * 6 * nq ** 3)  # host memory stands in for the device pointers: the call must fail before touching it
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    assert b.lib.b200fe_bk3_apply(p, nq, nelmt, ptr(basis), ptr(dbasis), ptr(buf), ptr(buf), ptr(buf), None) == 3
    assert b.lib.b200fe_bk1_apply(p, nq, nelmt, ptr(basis), ptr(buf), ptr(buf), ptr(buf), None) == 3
    assert b.lib.b200fe_sum_squares(8, ptr(buf), ptr(buf), None) == 3
    assert b.lib.b200fe_last_error() != b""
    # the operator object cannot even be created (its setup data live on the device)
    mesh = b.BoxMesh((1, 1, 1), 1, 2)
    with pytest.raises(Exception):
        b.LaplaceOperator(mesh)
