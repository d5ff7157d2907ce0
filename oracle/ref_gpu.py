"""oracle.ref_gpu -- the reference's OWN raw-CUDA kernels, compiled in place for sm_100a (oracle/Makefile ->
oracle/_ref/libref_gpu_ceedbk.so, libref_gpu_sumfact.so).  BENCH / TEST INFRASTRUCTURE ONLY: this is "the kernel
to beat"  timed on the same B200 beside ours, and a second parity witness (the
reference's GPU output on the same inputs).  Never imported by the product package.

  CEED_BK/include/kernels/BK{1,3,5}/templated_cuda_kernels.cuh, launch shape of
  CEED_BK/src/BK{1,3,5}/templated_cuda_benchmark.cc (main()), T = double.
  sum_factorization/include/kernels/BK1/templated_cuda_mma_kernels.cuh:123-232 (DMMA) and the CUDA-core
  warp-per-element kernel of the same study (.../BK1/templated_cuda_kernels.cuh:12-175).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_CEEDBK = os.path.join(_HERE, "_ref", "libref_gpu_ceedbk.so")
_SUMFACT = os.path.join(_HERE, "_ref", "libref_gpu_sumfact.so")
_libs: dict = {}


def available() -> bool:
    return os.path.exists(_CEEDBK) and os.path.exists(_SUMFACT)


def _lib(path):
    if path not in _libs:
        L = C.CDLL(path)
        vp = C.c_void_p
        if hasattr(L, "ref_gpu_ceedbk"):
            L.ref_gpu_ceedbk.restype = C.c_int
            L.ref_gpu_ceedbk.argtypes = [C.c_int, C.c_int, C.c_uint, vp, vp, vp, vp, vp, C.c_int, C.POINTER(C.c_float),
                                         C.POINTER(C.c_float), C.POINTER(C.c_uint)]
        if hasattr(L, "ref_gpu_sumfact_bk1"):
            L.ref_gpu_sumfact_bk1.restype = C.c_int
            L.ref_gpu_sumfact_bk1.argtypes = [C.c_int, C.c_int, C.c_uint, vp, vp, vp, vp, C.c_int, C.POINTER(C.c_float),
                                              C.POINTER(C.c_float), C.POINTER(C.c_uint)]
        _libs[path] = L
    return _libs[path]


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def ceedbk(kind: int, nq: int, nelmt: int, d_basis, d_dbasis, d_geom, d_in, d_out, ntests: int = 10):
    """Launches the reference kernel ntests (+2 warm-up) times on the current device's default stream.
    All arrays are float64 CUDA tensors.  Returns dict(ms_min, ms_mean, nelmtPerBatch, numBlocks, threads)."""
    import torch
    torch.cuda.synchronize()
    mn, me, shape = C.c_float(), C.c_float(), (C.c_uint * 3)()
    rc = _lib(_CEEDBK).ref_gpu_ceedbk(kind, nq, nelmt, _ptr(d_basis), _ptr(d_dbasis), _ptr(d_geom), _ptr(d_in), _ptr(d_out), ntests,
                                      C.byref(mn), C.byref(me), shape)
    if rc != 0:
        raise RuntimeError(f"reference CUDA kernel BK{kind} nq={nq}: cudaError {rc}")
    return dict(ms_min=mn.value, ms_mean=me.value, nelmtPerBatch=shape[0], numBlocks=shape[1], threads=shape[2])


def sumfact_bk1(variant: int, nq: int, nelmt: int, d_basis, d_JxW, d_in, d_out, ntests: int = 10):
    """variant 0: DMMA kernel (mma.sync.m8n8k4.f64), 1: CUDA-core warp-per-element kernel.  nq in (4, 8)."""
    import torch
    torch.cuda.synchronize()
    mn, me, shape = C.c_float(), C.c_float(), (C.c_uint * 3)()
    rc = _lib(_SUMFACT).ref_gpu_sumfact_bk1(variant, nq, nelmt, _ptr(d_basis), _ptr(d_JxW), _ptr(d_in), _ptr(d_out), ntests,
                                            C.byref(mn), C.byref(me), shape)
    if rc != 0:
        raise RuntimeError(f"reference sum_factorization BK1 variant {variant} nq={nq}: cudaError {rc}")
    return dict(ms_min=mn.value, ms_mean=me.value, warpsPerBlock=shape[0], numBlocks=shape[1], threads=shape[2])


def kernel_to_beat(b, dofs: float = 1e7, degrees=range(1, 9), kinds=("bk1", "bk3", "bk5"), ntests: int = 10, check: bool = True):
    """Times the reference's CUDA kernels and the product's E-vector kernels (module `b` = benchmarks_b200, passed in by
    the caller so this file imports nothing of the product) on identical device arrays at BASELINE config C2 sizes.
    Returns rows {kind, p, nelmt, ref_ms, ref_gdofs, ours_ms, ours_gdofs, speedup, max_rel_diff}."""
    import numpy as np
    import torch
    rows = []
    for kind in kinds:
        k = {"bk1": 1, "bk3": 3, "bk5": 5}[kind]
        for p in degrees:
            nm = p + 1
            nq = nm if k == 5 else p + 2
            nelmt = int(dofs) // nm ** 3
            basis = np.cos(np.arange(nq * nm, dtype=np.float64))
            dbasis = np.cos(np.arange(nq * nq, dtype=np.float64))
            g = torch.Generator(device="cuda").manual_seed(1000 * k + p)
            u = torch.rand(nelmt * nm ** 3, dtype=torch.float64, device="cuda", generator=g)
            geom = torch.rand(nelmt * (1 if k == 1 else 6) * nq ** 3, dtype=torch.float64, device="cuda", generator=g)
            d_basis, d_dbasis = torch.from_numpy(basis).cuda(), torch.from_numpy(dbasis).cuda()
            out_ref = torch.zeros_like(u)
            out = torch.empty_like(u)
            # the CUDA BK1 kernel indexes JxW as [r][q][p] (templated_cuda_kernels.cuh:111), the serial kernel (and ours) as
            # [p][q][r] (serial_kernels.hpp:77): hand the reference the transposed copy of the same field
            geom_ref = geom.view(nelmt, nq, nq, nq).permute(0, 3, 2, 1).contiguous() if k == 1 else geom
            r = ceedbk(k, nq, nelmt, d_basis, d_dbasis, geom_ref, u, out_ref, ntests)
            del geom_ref
            if k == 1:
                f = lambda: b.bk1_apply(p, nq, basis, geom, u, out)
            elif k == 3:
                f = lambda: b.bk3_apply(p, nq, basis, dbasis, geom, u, out)
            else:
                f = lambda: b.bk5_apply(p, dbasis, geom, u, out)
            for _ in range(3):
                f()
            torch.cuda.synchronize()
            ts = []
            for _ in range(ntests):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); f(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ours_min, ours_mean = float(min(ts)), float(np.mean(ts))
            ndof = nelmt * nm ** 3
            row = dict(kind=kind, p=p, nelmt=nelmt, n_dofs=ndof, ref_ms=r["ms_min"], ref_ms_mean=r["ms_mean"],
                       ref_gdofs=1e-6 * ndof / r["ms_min"], ours_ms=ours_min, ours_ms_mean=ours_mean, ours_gdofs=1e-6 * ndof / ours_min,
                       speedup=r["ms_min"] / ours_min, ref_launch=[r["nelmtPerBatch"], r["numBlocks"], r["threads"]])
            if check:  # (the CUDA BK5 kernel uses all six G components, templated_cuda_kernels.cuh:72-77 -- unlike the serial one)
                row["max_rel_diff"] = float((out - out_ref).abs().max() / out_ref.abs().max())
            rows.append(row)
            del u, geom, out, out_ref
            torch.cuda.empty_cache()
    return rows


def dmma_study(b, nelmt: int = 1 << 19, ntests: int = 10):
    """the reference's FP64 tensor-core BK1 kernel (nq = 4, its only instantiation; nq = 8 compiled from the same
    template) against its CUDA-core twin and against the product's BK1 kernel, same sizes.  GDoF/s = nelmt nm^3 / t."""
    import numpy as np
    import torch
    rows = []
    for nq in (4, 8):
        nm, p = nq - 1, nq - 2
        ne = nelmt if nq == 4 else nelmt // 8
        basis = np.cos(np.arange(nq * nm, dtype=np.float64))
        d_basis = torch.from_numpy(basis).cuda()
        g = torch.Generator(device="cuda").manual_seed(77 + nq)
        u = torch.rand(ne * nm ** 3, dtype=torch.float64, device="cuda", generator=g)
        JxW = torch.rand(ne * nq ** 3, dtype=torch.float64, device="cuda", generator=g)
        o0, o1, o2 = torch.zeros_like(u), torch.zeros_like(u), torch.empty_like(u)
        JxW_t = JxW.view(ne, nq, nq, nq).permute(0, 3, 2, 1).contiguous()  # these kernels index JxW as [r][q][p] (mma kernel :196)
        r0 = sumfact_bk1(0, nq, ne, d_basis, JxW_t, u, o0, ntests)
        r1 = sumfact_bk1(1, nq, ne, d_basis, JxW_t, u, o1, ntests)
        for _ in range(3):
            b.bk1_apply(p, nq, basis, JxW, u, o2)
        torch.cuda.synchronize()
        ts = []
        for _ in range(ntests):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); b.bk1_apply(p, nq, basis, JxW, u, o2); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ndof = ne * nm ** 3
        rows.append(dict(nq=nq, p=p, nelmt=ne, n_dofs=ndof,
                         dmma_ms=r0["ms_min"], dmma_gdofs=1e-6 * ndof / r0["ms_min"],
                         cuda_core_warp_ms=r1["ms_min"], cuda_core_warp_gdofs=1e-6 * ndof / r1["ms_min"],
                         ours_ms=float(min(ts)), ours_gdofs=1e-6 * ndof / float(min(ts)),
                         dmma_vs_cuda_core_max_rel_diff=float((o0 - o1).abs().max() / o1.abs().max()),
                         dmma_vs_ours_max_rel_diff=float((o0 - o2).abs().max() / o2.abs().max())))
        del u, JxW, JxW_t, o0, o1, o2
        torch.cuda.empty_cache()
    return rows
