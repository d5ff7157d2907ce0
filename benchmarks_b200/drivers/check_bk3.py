#!/usr/bin/env python
"""check_bk3.py -- the reference's CPU BK3 program (bk3_dealii/check_bk3.cc) on the B200 E-vector kernel.

Same problem: a strip of n hexahedra [0,n]x[0,1]x[0,1] whose lower-left vertices are lifted by
0.2 sin(2 pi x / L) (check_bk3.cc:32-52, "to avoid the Cartesian/affine cell optimisation"), MappingQ1,
FE_DGQ(p) element vectors, QGauss(p+2), input in[i] = 0.23 + 0.12 sin(pi i/(N-1)) - 0.02 sin(52 pi i/(N-1))
(:72-74), operation evaluate(gradients) / submit_gradient / integrate(gradients) = the BK3 Laplacian (:86-113);
10 repetitions, best / average / worst, GDoF/s = n_dofs / best (:115-139).  FP64.

    python benchmarks_b200/drivers/check_bk3.py [degree=3] [max_elements=1000000]
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import benchmarks_b200 as b  # noqa: E402
from benchmarks_b200._lib import check, lib  # noqa: E402


def strip_nodes(n_elements: int, device="cuda") -> torch.Tensor:
    """Vertices of the deformed strip as MappingQ1 support points: nodes[cell][3][c][b][a] (a <-> x)."""
    e = torch.arange(n_elements, dtype=torch.float64, device=device)
    length = float(n_elements)
    nodes = torch.empty(n_elements, 3, 2, 2, 2, dtype=torch.float64, device=device)
    for c in range(2):
        for bb in range(2):
            for a in range(2):
                x = e + a
                y = torch.full_like(e, float(bb))
                if bb == 0 and c == 0:  # vertices on the line y = z = 0 that are vertex 0 of some cell (x < n)
                    y = torch.where(x < n_elements, 0.2 * torch.sin(2.0 * x * np.pi / length), y)
                nodes[:, 0, c, bb, a] = x
                nodes[:, 1, c, bb, a] = y
                nodes[:, 2, c, bb, a] = float(c)
    return nodes.contiguous()


def setup(degree: int, n_elements: int):
    nq, nm = degree + 2, degree + 1
    bas = b.basis_1d(degree, nq, b.QUAD_GAUSS)
    basis = np.ascontiguousarray(bas["shape_values"].reshape(nm, nq).T)        # basis[q*nm+i]
    dbasis = np.ascontiguousarray(bas["co_shape_gradients"].reshape(nq, nq).T)  # dbasis[p*nq+n]
    nodes = strip_nodes(n_elements)
    G = torch.empty(n_elements * 6 * nq ** 3, dtype=torch.float64, device="cuda")
    check(lib.b200fe_geometry_from_nodes(1, nq, b.QUAD_GAUSS, n_elements, C.c_void_p(nodes.data_ptr()), C.c_void_p(G.data_ptr()), None,
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    n = n_elements * nm ** 3
    i = torch.arange(n, dtype=torch.float64, device="cuda")
    u = 0.23 + 0.12 * torch.sin(np.pi * i / (n - 1)) - 0.02 * torch.sin(52.0 * np.pi * i / (n - 1))
    return basis, dbasis, nodes, G, u


def test_bk(degree: int, n_elements: int, n_tests: int, print_header: bool):
    basis, dbasis, _, G, u = setup(degree, n_elements)
    out = torch.empty_like(u)
    times = []
    for _ in range(n_tests):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        b.bk3_apply(degree, degree + 2, basis, dbasis, G, u, out)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) * 1e-3)
    best, avg, worst = min(times), sum(times) / len(times), max(times)
    if print_header:
        print("test in FP64\n  p  |  q  |     n_dofs |    min_t |    avg_t |    max_t |   GDoF/s")
    print(f" {degree:2d}  | {degree + 2:2d}  | {u.numel():10d} | {best:8.2e} | {avg:8.2e} | {worst:8.2e} | {1e-9 * u.numel() / best:8.3f}")


def main():
    degree = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    max_elements = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
    size = 1000
    while size < max_elements:
        test_bk(degree, size, 10, size == 1000)
        size *= 2


if __name__ == "__main__":
    main()
