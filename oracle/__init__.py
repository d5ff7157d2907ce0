"""oracle -- CPU checker for the b200fe hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (benchmarks_b200/) never does.

  oracle.port   ctypes view of liboracle.so   (our C restatement: bk_oracle.c, fe_oracle.c)
  oracle.ref    ctypes view of oracle/_ref/*.so (the reference's own serial kernels compiled in
                place from /root/reference by oracle/Makefile); None when not built
  oracle.fe     numpy restatement of bases / mesh / numbering / operator / CG (fe_oracle.py)
  oracle.hanging  two-level mesh with hanging-node constraints (hanging_oracle.py)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import fe_oracle as fe  # noqa: F401
from . import hanging_oracle as hanging  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_up = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")


def build(quiet: bool = True) -> None:
    """make -C oracle (liboracle.so always; _ref only when /root/reference exists)."""
    subprocess.run(["make", "-C", _HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


class _Port:
    def __init__(self):
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = self.lib = C.CDLL(path)
        L.oracle_bk1.restype = C.c_double
        L.oracle_bk1.argtypes = [C.c_int, C.c_int, C.c_uint, _dp, _dp, _dp, _dp]
        L.oracle_bk1_direct.restype = C.c_double
        L.oracle_bk1_direct.argtypes = L.oracle_bk1.argtypes
        L.oracle_bk3.restype = C.c_double
        L.oracle_bk3.argtypes = [C.c_int, C.c_int, C.c_uint, _dp, _dp, _dp, C.c_int, _dp, _dp]
        L.oracle_bk5.restype = C.c_double
        L.oracle_bk5.argtypes = [C.c_int, C.c_uint, _dp, _dp, C.c_int, _dp, _dp]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_num_threads.restype = None
        L.oracle_set_num_threads.argtypes = [C.c_int]
        vp = C.c_void_p
        L.oracle_op_apply.restype = None
        L.oracle_op_apply.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32,
                                      _dp, _dp, vp, vp, _up, C.c_int, vp, vp, C.c_uint32, vp, _dp, _dp]
        L.oracle_cg_solve.restype = C.c_int
        L.oracle_cg_solve.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32,
                                      _dp, _dp, vp, vp, _up, C.c_int, vp, vp, C.c_uint32, vp,
                                      vp, _dp, _dp, C.c_int, C.c_double, C.c_double,
                                      C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]

    # ---- E-vector kernels -------------------------------------------------
    def bk1(self, nm, nq, basis, JxW, u, direct=False):
        nelmt = u.size // nm ** 3
        out = np.empty_like(u)
        f = self.lib.oracle_bk1_direct if direct else self.lib.oracle_bk1
        s = f(nm, nq, nelmt, np.ascontiguousarray(basis), np.ascontiguousarray(JxW), np.ascontiguousarray(u), out)
        return out, s

    def bk3(self, nm, nq, basis, dbasis, G, u, g_layout=1):
        nelmt = u.size // nm ** 3
        out = np.empty_like(u)
        s = self.lib.oracle_bk3(nm, nq, nelmt, np.ascontiguousarray(basis), np.ascontiguousarray(dbasis),
                                np.ascontiguousarray(G), g_layout, np.ascontiguousarray(u), out)
        return out, s

    def bk5(self, nq, dbasis, G, u, g_layout=1):
        nelmt = u.size // nq ** 3
        out = np.empty_like(u)
        s = self.lib.oracle_bk5(nq, nelmt, np.ascontiguousarray(dbasis), np.ascontiguousarray(G), g_layout,
                                np.ascontiguousarray(u), out)
        return out, s

    # ---- L-vector operator + CG ------------------------------------------
    @staticmethod
    def _opt(a, dtype):
        if a is None:
            return None, None
        a = np.ascontiguousarray(a, dtype=dtype)
        return a, a.ctypes.data_as(C.c_void_p)

    def _common(self, nm, nq, collocated, flags, shape_values, co_shape_gradients, G, JxW, dof_indices,
                colors, constrained, n_local):
        keep = []
        Gk, Gp = self._opt(G, np.float64)
        Jk, Jp = self._opt(JxW, np.float64)
        idx = np.ascontiguousarray(dof_indices, dtype=np.uint32)
        n_cells = idx.size // nm ** 3
        if colors is not None:
            off, cells = colors
            offk, offp = self._opt(off, np.uint32)
            ck, cp = self._opt(cells, np.uint32)
            ncol = len(offk) - 1
        else:
            offk = ck = offp = cp = None
            ncol = 0
        conk, conp = self._opt(constrained if constrained is not None else np.zeros(0, np.uint32), np.uint32)
        keep += [Gk, Jk, idx, offk, ck, conk]
        args = [nm, nq, int(collocated), int(flags), n_cells, int(n_local),
                np.ascontiguousarray(shape_values, dtype=np.float64),
                np.ascontiguousarray(co_shape_gradients, dtype=np.float64), Gp, Jp, idx, ncol, offp, cp,
                len(conk), conp]
        return args, keep

    def op_apply(self, src, *, nm, nq, collocated, flags, shape_values, co_shape_gradients, G, JxW,
                 dof_indices, colors=None, constrained=None):
        args, keep = self._common(nm, nq, collocated, flags, shape_values, co_shape_gradients, G, JxW,
                                  dof_indices, colors, constrained, len(src))
        dst = np.empty_like(src)
        self.lib.oracle_op_apply(*args, np.ascontiguousarray(src), dst)
        return dst

    def cg_solve(self, b, *, nm, nq, collocated, flags, shape_values, co_shape_gradients, G, JxW,
                 dof_indices, colors=None, constrained=None, inv_diag=None, max_it=1000, abs_tol=1e-16,
                 rel_tol=1e-9):
        args, keep = self._common(nm, nq, collocated, flags, shape_values, co_shape_gradients, G, JxW,
                                  dof_indices, colors, constrained, len(b))
        dk, dp = self._opt(inv_diag, np.float64)
        x = np.empty_like(b)
        r0, rn, ok = C.c_double(), C.c_double(), C.c_int()
        its = self.lib.oracle_cg_solve(*args, dp, np.ascontiguousarray(b), x, int(max_it), float(abs_tol),
                                       float(rel_tol), C.byref(r0), C.byref(rn), C.byref(ok))
        return x, its, r0.value, rn.value, bool(ok.value)

    def num_threads(self):
        return self.lib.oracle_num_threads()

    def set_num_threads(self, n: int):
        """OpenMP team size of the timed CPU arm (torchrun exports OMP_NUM_THREADS=1 to its ranks)."""
        self.lib.oracle_set_num_threads(int(n))


class _Ref:
    """The reference's serial kernels (nm = nq-1 for BK1/BK3 by construction)."""

    def __init__(self, ceedbk, sumfact):
        self.c = C.CDLL(ceedbk)
        self.s = C.CDLL(sumfact)
        for f in (self.c.ref_ceedbk_bk1, self.s.ref_sumfact_bk1, self.s.ref_sumfact_bk1_direct):
            f.restype = C.c_double
            f.argtypes = [C.c_uint, C.c_uint, _dp, _dp, _dp, _dp]
        self.c.ref_ceedbk_bk3.restype = C.c_double
        self.c.ref_ceedbk_bk3.argtypes = [C.c_uint, C.c_uint, _dp, _dp, _dp, _dp, _dp]
        for f in (self.c.ref_ceedbk_bk5, self.s.ref_sumfact_bk5):
            f.restype = C.c_double
            f.argtypes = [C.c_uint, C.c_uint, _dp, _dp, _dp, _dp]

    def bk1(self, nq, basis, JxW, u, which="ceedbk"):
        f = {"ceedbk": self.c.ref_ceedbk_bk1, "sumfact": self.s.ref_sumfact_bk1,
             "direct": self.s.ref_sumfact_bk1_direct}[which]
        nelmt = u.size // (nq - 1) ** 3
        out = np.zeros_like(u)
        s = f(nq, nelmt, np.ascontiguousarray(basis), np.ascontiguousarray(JxW), u.copy(), out)
        return out, s

    def bk3(self, nq, basis, dbasis, G_serial_layout, u):
        """G in the serial layout [e][p][q][6][r] (BK3 serial_kernels.hpp:87-92)."""
        nelmt = u.size // (nq - 1) ** 3
        out = np.zeros_like(u)
        s = self.c.ref_ceedbk_bk3(nq, nelmt, np.ascontiguousarray(basis), np.ascontiguousarray(dbasis),
                                  np.ascontiguousarray(G_serial_layout), u.copy(), out)
        return out, s

    def bk5(self, nq, dbasis, G, u, which="sumfact"):
        """which='sumfact': G [e][i][j][6][k]; which='ceedbk': G [e][6][i][j][k] but only comp. 0 read."""
        f = self.s.ref_sumfact_bk5 if which == "sumfact" else self.c.ref_ceedbk_bk5
        nelmt = u.size // nq ** 3
        out = np.zeros_like(u)
        s = f(nq, nelmt, np.ascontiguousarray(dbasis), np.ascontiguousarray(G), np.ascontiguousarray(u), out)
        return out, s


def _load_ref():
    a = os.path.join(_HERE, "_ref", "libref_ceedbk.so")
    b = os.path.join(_HERE, "_ref", "libref_sumfact.so")
    if os.path.exists(a) and os.path.exists(b):
        return _Ref(a, b)
    return None


port = _Port()
ref = _load_ref()


# ---- the reference's seedless synthetic inputs (CEED_BK/src/BK3/templated_cuda_benchmark.cc:45-66)
def kat_inputs(kind: str, p: int, nelmt: int, nq: int | None = None):
    """in = 3, JxW = 1, G = 2, basis[q*nm+i] = cos(q*nm+i), dbasis[i*nq+n] = cos(i*nq+n)."""
    if kind == "bk5":
        nq = p + 1 if nq is None else nq
        nm = nq
    else:
        nq = p + 2 if nq is None else nq
        nm = p + 1
    basis = np.cos(np.arange(nq * nm, dtype=np.float64))
    dbasis = np.cos(np.arange(nq * nq, dtype=np.float64))
    u = np.full(nelmt * nm ** 3, 3.0)
    JxW = np.ones(nelmt * nq ** 3)
    G = np.full(nelmt * 6 * nq ** 3, 2.0)
    return dict(nm=nm, nq=nq, basis=basis, dbasis=dbasis, u=u, JxW=JxW, G=G)
