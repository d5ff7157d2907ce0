"""Element-vector bake-off kernels BK1 / BK3 / BK5 (host mirror of the reference's standalone
drivers, CEED_BK/src/BK{1,3,5}/templated_cuda_benchmark.cc: same arrays, same layouts)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import check, lib


def _stream_ptr(stream=None):
    s = torch.cuda.current_stream() if stream is None else stream
    return C.c_void_p(s.cuda_stream)


def _dev(t: torch.Tensor, name: str):
    if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
        raise ValueError(f"{name} must be a contiguous float64 CUDA tensor")
    return C.c_void_p(t.data_ptr())


def _host(a, n: int, name: str):
    a = np.ascontiguousarray(a, dtype=np.float64).ravel()
    if a.size != n:
        raise ValueError(f"{name} must hold {n} doubles, got {a.size}")
    return a


def bk1_apply(p: int, nq: int, basis, JxW: torch.Tensor, u: torch.Tensor, out: torch.Tensor | None = None, stream=None):
    """out_e = B^T (JxW .* (B u_e)); basis[q*nm+i] on the host, the rest on the device."""
    nm = p + 1
    nelmt = u.numel() // nm ** 3
    b = _host(basis, nq * nm, "basis")
    out = torch.empty_like(u) if out is None else out
    check(lib.b200fe_bk1_apply(p, nq, nelmt, b.ctypes.data_as(C.c_void_p), _dev(JxW, "JxW"), _dev(u, "u"),
                               _dev(out, "out"), _stream_ptr(stream)))
    return out


def bk3_apply(p: int, nq: int, basis, dbasis, G: torch.Tensor, u: torch.Tensor, out: torch.Tensor | None = None, stream=None):
    """out_e = B^T D^T G D B u_e; G [e][6][nq^3]."""
    nm = p + 1
    nelmt = u.numel() // nm ** 3
    b = _host(basis, nq * nm, "basis")
    d = _host(dbasis, nq * nq, "dbasis")
    out = torch.empty_like(u) if out is None else out
    check(lib.b200fe_bk3_apply(p, nq, nelmt, b.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p),
                               _dev(G, "G"), _dev(u, "u"), _dev(out, "out"), _stream_ptr(stream)))
    return out


def bk5_apply(p: int, dbasis, G: torch.Tensor, u: torch.Tensor, out: torch.Tensor | None = None, stream=None):
    """out_e = D^T G D u_e with nm = nq = p+1."""
    nq = p + 1
    nelmt = u.numel() // nq ** 3
    d = _host(dbasis, nq * nq, "dbasis")
    out = torch.empty_like(u) if out is None else out
    check(lib.b200fe_bk5_apply(p, nelmt, d.ctypes.data_as(C.c_void_p), _dev(G, "G"), _dev(u, "u"),
                               _dev(out, "out"), _stream_ptr(stream)))
    return out


def sum_squares(x: torch.Tensor, stream=None) -> torch.Tensor:
    """Device scalar sum(x^2): the drivers' `check` column is its square root."""
    res = torch.empty(1, dtype=torch.float64, device=x.device)
    check(lib.b200fe_sum_squares(x.numel(), _dev(x, "x"), _dev(res, "res"), _stream_ptr(stream)))
    return res


def bk_launch_info(kind: int, p: int, nq: int, nelmt: int):
    e, g, t, s = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    check(lib.b200fe_bk_launch_info(kind, p, nq, nelmt, C.byref(e), C.byref(g), C.byref(t), C.byref(s)))
    return dict(elems_per_block=e.value, num_blocks=g.value, threads_per_block=t.value, smem_bytes=s.value)
