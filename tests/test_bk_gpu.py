"""GPU parity of the element-vector kernels BK1/BK3/BK5 through the C ABI against the CPU oracle.

Tolerance: FP64 results agree with the oracle to <= 1e-12 relative max-norm (north star); the
reference's known-answer norms (synthetic seedless inputs, sqrt(sum out^2)) to 1e-12 relative.
"""
import json
import os

import numpy as np
import pytest
import torch

from tests.golden.make_bk_golden import random_case

pytestmark = pytest.mark.gpu
TOL = 1e-12


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rel_max(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def gold(golden_dir):
    with open(os.path.join(golden_dir, "bk_norms.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("p", range(1, 9))
def test_kat_norms(oracle_mod, gold, p):
    """The reference's own check: in=3, JxW=1, G=2, basis=cos(idx) -> sqrt(sum out^2)."""
    import benchmarks_b200 as b
    g = gold["kat_norms_nelmt64"][str(p)]
    k = oracle_mod.kat_inputs("bk1", p, 64)
    o1 = b.bk1_apply(p, k["nq"], k["basis"], dev(k["JxW"]), dev(k["u"]))
    o3 = b.bk3_apply(p, k["nq"], k["basis"], k["dbasis"], dev(k["G"]), dev(k["u"]))
    k5 = oracle_mod.kat_inputs("bk5", p, 64)
    o5 = b.bk5_apply(p, k5["dbasis"], dev(k5["G"]), dev(k5["u"]))
    for out, key in ((o1, "bk1"), (o3, "bk3"), (o5, "bk5")):
        n = float(torch.sqrt(b.sum_squares(out)).item())
        assert n == pytest.approx(g[key], rel=TOL), key
        assert n == pytest.approx(float(np.sqrt((out.cpu().numpy() ** 2).sum())), rel=1e-13)


@pytest.mark.parametrize("p", range(1, 9))
@pytest.mark.parametrize("nelmt", [1, 37, 1000])
def test_random_inputs_elementwise(oracle_mod, p, nelmt):
    """Random B, D, G, JxW, u: catches index permutations the constant-input KATs cannot."""
    import benchmarks_b200 as b
    c = random_case("bk1", p, nelmt, 10 * p + nelmt)
    ref1, _ = oracle_mod.port.bk1(c["nm"], c["nq"], c["basis"], c["JxW"], c["u"])
    ref3, _ = oracle_mod.port.bk3(c["nm"], c["nq"], c["basis"], c["dbasis"], c["G"].ravel(), c["u"], 1)
    o1 = b.bk1_apply(p, c["nq"], c["basis"], dev(c["JxW"]), dev(c["u"])).cpu().numpy()
    o3 = b.bk3_apply(p, c["nq"], c["basis"], c["dbasis"], dev(c["G"].ravel()), dev(c["u"])).cpu().numpy()
    assert rel_max(o1, ref1) <= TOL
    assert rel_max(o3, ref3) <= TOL
    c5 = random_case("bk5", p, nelmt, 20 * p + nelmt)
    ref5, _ = oracle_mod.port.bk5(c5["nq"], c5["dbasis"], c5["G"].ravel(), c5["u"], 1)
    o5 = b.bk5_apply(p, c5["dbasis"], dev(c5["G"].ravel()), dev(c5["u"])).cpu().numpy()
    assert rel_max(o5, ref5) <= TOL


def test_empty_input_is_a_noop():
    import benchmarks_b200 as b
    z = torch.empty(0, dtype=torch.float64, device="cuda")
    out = b.bk5_apply(3, np.zeros(16), z, z)
    assert out.numel() == 0
    assert float(b.sum_squares(z).item()) == 0.0


@pytest.mark.parametrize("kind,p", [("bk5", 6), ("bk3", 4), ("bk1", 8), ("bk3", 8)])
def test_full_size_properties(kind, p):
    """BASELINE config C2 size (~1e7 E-vector DoFs): linearity and symmetry of the element operator,
    size-independent properties checked without a CPU run: <A u, v> = <u, A v>, A(au+bv) = aAu+bAv."""
    import benchmarks_b200 as b
    nm = p + 1
    nq = nm if kind == "bk5" else p + 2
    nelmt = 10_000_000 // nm ** 3
    gen = torch.Generator(device="cuda").manual_seed(p)
    rnd = lambda n: torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
    basis = np.cos(np.arange(nq * nm, dtype=np.float64))
    dbasis = np.cos(np.arange(nq * nq, dtype=np.float64))
    u, v = rnd(nelmt * nm ** 3), rnd(nelmt * nm ** 3)
    if kind == "bk1":
        J = rnd(nelmt * nq ** 3) + 2
        A = lambda x: b.bk1_apply(p, nq, basis, J, x)
    else:
        G = rnd(nelmt * 6 * nq ** 3)
        A = (lambda x: b.bk5_apply(p, dbasis, G, x)) if kind == "bk5" else (lambda x: b.bk3_apply(p, nq, basis, dbasis, G, x))
    Au, Av = A(u), A(v)
    lhs, rhs = torch.dot(Au, v).item(), torch.dot(u, Av).item()
    scale = (Au.norm() * v.norm()).item()
    assert abs(lhs - rhs) <= 1e-12 * scale
    comb = A(0.5 * u - 1.25 * v)
    assert (comb - (0.5 * Au - 1.25 * Av)).abs().max().item() <= 1e-12 * Au.abs().max().item()


@pytest.mark.parametrize("degree", [3, 4, 5])
def test_check_bk3_deformed_strip_matches_oracle(oracle_mod, degree):
    """The reference's CPU BK3 program (bk3_dealii/check_bk3.cc): deformed strip, MappingQ1, FE_DGQ element vectors,
    QGauss(p+2), its input vector -- against the oracle (numpy geometry from the same vertices + serial BK3)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("check_bk3", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                            "benchmarks_b200", "drivers", "check_bk3.py"))
    drv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(drv)
    import benchmarks_b200 as b
    fe = oracle_mod.fe
    n = 24
    basis, dbasis, nodes, G, u = drv.setup(degree, n)
    out = b.bk3_apply(degree, degree + 2, basis, dbasis, G, u).cpu().numpy()
    bas = fe.basis_1d(degree, degree + 2)
    Go, _ = fe.geometric_factors(nodes.cpu().numpy(), 1, bas)
    assert np.abs(G.cpu().numpy().reshape(Go.shape) - Go).max() <= TOL * np.abs(Go).max()
    ref, _ = oracle_mod.port.bk3(degree + 1, degree + 2, bas["B"].ravel(), bas["D"].ravel(), Go.ravel(), u.cpu().numpy(), 1)
    assert rel_max(out, ref) <= TOL
    # the deformation is really there: first cell is not a cube (off-diagonal geometric factors present)
    assert np.abs(Go[0, 1]).max() > 1e-3 * np.abs(Go[0, 0]).max()


def test_matches_the_reference_cuda_kernels_elementwise():
    """The reference's OWN CUDA kernels (CEED_BK/include/kernels/BK{1,3,5}/templated_cuda_kernels.cuh, T = double, compiled in
    place into oracle/_ref/libref_gpu_ceedbk.so, the drivers' launch shape) on the same device arrays: every degree, <= 1e-13."""
    import benchmarks_b200 as b
    from oracle import ref_gpu
    if not ref_gpu.available():
        pytest.skip("oracle/_ref/libref_gpu_*.so not built in this checkout")
    rows = ref_gpu.kernel_to_beat(b, dofs=3e5, ntests=1)
    assert len(rows) == 24
    for r in rows:
        assert r["max_rel_diff"] <= 1e-13, (r["kind"], r["p"], r["max_rel_diff"])
