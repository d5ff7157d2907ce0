"""Pins the CPU oracle BEFORE it is trusted as the checker (no GPU needed).

 * C restatement (oracle/bk_oracle.c) vs the reference's own serial kernels compiled in place
   (oracle/_ref), bit for bit, on the reference's synthetic inputs and on random inputs;
 * vs the committed golden norms (tests/golden/bk_norms.json);
 * sum factorisation vs brute force (the reference's only cross-algorithm check,
   sum_factorization/tests/BK1/serial_verification.cc:49-55);
 * numpy FE oracle + C L-vector oracle vs the p=4 CG goldens of
   CEED_bp/results/1xGH200_P4.txt:636-640 and vs the element-free Kronecker operator.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from tests.golden.make_bk_golden import random_case, serial_G


@pytest.fixture(scope="module")
def gold(golden_dir):
    with open(os.path.join(golden_dir, "bk_norms.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("p", range(1, 9))
def test_port_matches_golden_norms(oracle_mod, gold, p):
    k = oracle_mod.kat_inputs("bk1", p, 64)
    _, s1 = oracle_mod.port.bk1(k["nm"], k["nq"], k["basis"], k["JxW"], k["u"])
    _, s3 = oracle_mod.port.bk3(k["nm"], k["nq"], k["basis"], k["dbasis"], k["G"], k["u"], 1)
    k5 = oracle_mod.kat_inputs("bk5", p, 64)
    _, s5 = oracle_mod.port.bk5(k5["nq"], k5["dbasis"], k5["G"], k5["u"], 1)
    g = gold["kat_norms_nelmt64"][str(p)]
    assert np.sqrt(s1) == pytest.approx(g["bk1"], rel=1e-13)
    assert np.sqrt(s3) == pytest.approx(g["bk3"], rel=1e-13)
    assert np.sqrt(s5) == pytest.approx(g["bk5"], rel=1e-13)


def test_port_matches_survey_table(oracle_mod):
    # the 8-digit values printed by the reference's serial_verification programs
    table = {1: (769.87113, 5033.9152, 960.20391), 4: (2329.6413, 115436.4, 2077.5619),
             8: (373871.48, 483268.65, 16895.147)}
    for p, (n1, n3, n5) in table.items():
        k = oracle_mod.kat_inputs("bk1", p, 64)
        assert np.sqrt(oracle_mod.port.bk1(k["nm"], k["nq"], k["basis"], k["JxW"], k["u"])[1]) == pytest.approx(n1, rel=6e-8)
        assert np.sqrt(oracle_mod.port.bk3(k["nm"], k["nq"], k["basis"], k["dbasis"], k["G"], k["u"])[1]) == pytest.approx(n3, rel=6e-8)
        k5 = oracle_mod.kat_inputs("bk5", p, 64)
        assert np.sqrt(oracle_mod.port.bk5(k5["nq"], k5["dbasis"], k5["G"], k5["u"])[1]) == pytest.approx(n5, rel=6e-8)


def test_nelmt1000_spot_values(oracle_mod, gold):
    k = oracle_mod.kat_inputs("bk1", 2, 1000)
    g = gold["kat_norms_nelmt1000"]
    assert np.sqrt(oracle_mod.port.bk1(k["nm"], k["nq"], k["basis"], k["JxW"], k["u"])[1]) == pytest.approx(g["bk1_p2"], rel=1e-13)
    assert np.sqrt(oracle_mod.port.bk1(k["nm"], k["nq"], k["basis"], k["JxW"], k["u"], direct=True)[1]) == pytest.approx(g["bk1_p2_direct"], rel=1e-12)
    assert g["bk1_p2"] == pytest.approx(32890.7, rel=1e-6)
    assert np.sqrt(oracle_mod.port.bk3(k["nm"], k["nq"], k["basis"], k["dbasis"], k["G"], k["u"])[1]) == pytest.approx(113018.16, rel=1e-7)


@pytest.mark.parametrize("p", range(1, 9))
def test_port_random_inputs_bitexact_vs_reference_digest(oracle_mod, gold, p):
    """sha256 of the reference's output vectors on seeded random inputs (works without _ref)."""
    c = random_case("bk1", p, 5, 1000 + p)
    o1, _ = oracle_mod.port.bk1(c["nm"], c["nq"], c["basis"], c["JxW"], c["u"])
    o3, _ = oracle_mod.port.bk3(c["nm"], c["nq"], c["basis"], c["dbasis"], c["G"].ravel(), c["u"], 1)
    o3s, _ = oracle_mod.port.bk3(c["nm"], c["nq"], c["basis"], c["dbasis"], serial_G(c["G"]).ravel(), c["u"], 0)
    c5 = random_case("bk5", p, 5, 2000 + p)
    o5, _ = oracle_mod.port.bk5(c5["nq"], c5["dbasis"], c5["G"].ravel(), c5["u"], 1)
    g = gold["random_sha256"][str(p)]
    assert hashlib.sha256(o1.tobytes()).hexdigest() == g["bk1"]
    assert hashlib.sha256(o3.tobytes()).hexdigest() == g["bk3"]
    assert hashlib.sha256(o3s.tobytes()).hexdigest() == g["bk3"]
    assert hashlib.sha256(o5.tobytes()).hexdigest() == g["bk5"]


@pytest.mark.parametrize("p", [1, 3, 6, 8])
def test_port_bitexact_vs_compiled_reference(oracle_mod, p):
    if oracle_mod.ref is None:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    c = random_case("bk1", p, 3, 77 + p)
    r1, _ = oracle_mod.ref.bk1(c["nq"], c["basis"], c["JxW"], c["u"])
    o1, _ = oracle_mod.port.bk1(c["nm"], c["nq"], c["basis"], c["JxW"], c["u"])
    assert np.array_equal(r1, o1)
    r3, _ = oracle_mod.ref.bk3(c["nq"], c["basis"], c["dbasis"], serial_G(c["G"]).ravel(), c["u"])
    o3, _ = oracle_mod.port.bk3(c["nm"], c["nq"], c["basis"], c["dbasis"], c["G"].ravel(), c["u"], 1)
    assert np.array_equal(r3, o3)
    c5 = random_case("bk5", p, 3, 99 + p)
    r5, _ = oracle_mod.ref.bk5(c5["nq"], c5["dbasis"], serial_G(c5["G"]).ravel(), c5["u"])
    o5, _ = oracle_mod.port.bk5(c5["nq"], c5["dbasis"], c5["G"].ravel(), c5["u"], 1)
    assert np.array_equal(r5, o5)
    # CEED_BK's BK5 serial kernel reads component 0 for all six factors: only equal for constant G
    k5 = oracle_mod.kat_inputs("bk5", p, 3)
    rq, _ = oracle_mod.ref.bk5(k5["nq"], k5["dbasis"], k5["G"], k5["u"], which="ceedbk")
    oq, _ = oracle_mod.port.bk5(k5["nq"], k5["dbasis"], k5["G"], k5["u"], 1)
    assert np.array_equal(rq, oq)


@pytest.mark.parametrize("p", [1, 2, 4])
def test_sum_factorisation_equals_brute_force(oracle_mod, p):
    c = random_case("bk1", p, 4, 5 + p)
    a, _ = oracle_mod.port.bk1(c["nm"], c["nq"], c["basis"], c["JxW"], c["u"])
    b, _ = oracle_mod.port.bk1(c["nm"], c["nq"], c["basis"], c["JxW"], c["u"], direct=True)
    assert np.abs(a - b).max() <= 1e-13 * np.abs(b).max()


# ---------------------------------------------------------------------------------------------
# L-vector path: bases, numbering, operator, CG
# ---------------------------------------------------------------------------------------------
def test_1d_bases_partition_of_unity_and_exactness(oracle_mod):
    fe = oracle_mod.fe
    for p in range(1, 9):
        for nq, quad in ((p + 2, "gauss"), (p + 1, "gauss"), (p + 1, "gll")):
            b = fe.basis_1d(p, nq, quad)
            assert np.allclose(b["B"].sum(axis=1), 1.0, atol=1e-13)      # sum_i phi_i = 1 (test_lagrange.cc idea)
            assert np.allclose(b["Bg"].sum(axis=1), 0.0, atol=1e-10)
            assert np.allclose(b["D"].sum(axis=1), 0.0, atol=1e-10)
            assert b["wq"].sum() == pytest.approx(1.0, abs=1e-14)
            # D differentiates polynomials of degree < nq exactly at the points
            x = b["xq"]
            assert np.allclose(b["D"] @ x ** (nq - 1), (nq - 1) * x ** (nq - 2), atol=1e-9)


def test_fe_q_numbering_counts(oracle_mod):
    fe = oracle_mod.fe
    for p in (1, 2, 3):
        mesh = fe.BoxMesh((2, 1, 1), 1)
        d = fe.distribute_dofs(mesh, p, 1)
        assert len(d["lattice_of_global"]) == np.prod([c * p + 1 for c in mesh.cells])
        h2l = fe.hierarchic_to_lexicographic(p)
        assert sorted(h2l.tolist()) == list(range((p + 1) ** 3))
    # first cell of a Q2 mesh: vertices first, then lines, quads, interior (A2/A3)
    mesh = fe.BoxMesh((1, 1, 1), 1)
    d = fe.distribute_dofs(mesh, 2, 1)
    first = fe.cell_dofs_global(mesh, d, 0)  # lexicographic view of cell 0
    h2l = fe.hierarchic_to_lexicographic(2)
    assert first[h2l].tolist() == list(range(27))


@pytest.fixture(scope="module")
def cg_golden(golden_dir):
    with open(os.path.join(golden_dir, "bp3_cg_p4.json")) as f:
        return json.load(f)["rows"]


def _bp3_setup(fe, cycle, p, nq, quad="gauss"):
    mesh = fe.BoxMesh.bp3_cycle(cycle)
    dofs = fe.distribute_dofs(mesh, p, 1)
    rd = fe.rank_data(mesh, dofs, 0)
    bas = fe.basis_1d(p, nq, quad)
    G, JxW = fe.geometric_factors(fe.cell_nodes(mesh, rd["cells"], 1), 1, bas)
    return mesh, dofs, rd, bas, G, JxW


def _colors(mesh, cells):
    xyz = mesh.cell_xyz[cells]
    col = (xyz[:, 0] & 1) + 2 * (xyz[:, 1] & 1) + 4 * (xyz[:, 2] & 1)
    order = np.argsort(col, kind="stable")
    off = np.concatenate([[0], np.cumsum(np.bincount(col, minlength=8))])
    return off.astype(np.uint32), order.astype(np.uint32)


@pytest.mark.parametrize("nq", [6, 5])
def test_cg_iteration_goldens_numpy(oracle_mod, cg_golden, nq):
    """bp3 protocol (bp3.cc:266-285): rhs = int phi, x0 = 0, rel tol 1e-9 -> 92 its, rate 0.7959."""
    fe = oracle_mod.fe
    cycle, cells, ndofs, its_g, red_g = cg_golden[0]
    mesh, dofs, rd, bas, G, JxW = _bp3_setup(fe, cycle, 4, nq)
    assert mesh.n_cells == cells and rd["n_owned"] == ndofs
    b = fe.rhs_one(rd, bas, JxW)
    x, its, r0, rn, ok = fe.solver_cg(lambda v: fe.op_apply(v, rd, bas, G), b, 10 ** 9, 1e-16, 1e-9)
    assert ok and its == its_g
    assert (rn / r0) ** (1.0 / its) == pytest.approx(red_g, abs=5e-5)


@pytest.mark.parametrize("row", [0, 1, 2])
def test_cg_iteration_goldens_c_oracle(oracle_mod, cg_golden, row):
    fe = oracle_mod.fe
    cycle, cells, ndofs, its_g, red_g = cg_golden[row]
    mesh, dofs, rd, bas, G, JxW = _bp3_setup(fe, cycle, 4, 6)
    assert mesh.n_cells == cells and rd["n_owned"] == ndofs
    b = fe.rhs_one(rd, bas, JxW)
    x, its, r0, rn, ok = oracle_mod.port.cg_solve(
        b, nm=5, nq=6, collocated=False, flags=1, shape_values=bas["B"].T.copy(),
        co_shape_gradients=bas["D"].T.copy(), G=G, JxW=None, dof_indices=rd["dof_indices"],
        colors=_colors(mesh, rd["cells"]), constrained=rd["constrained"], max_it=10 ** 9, rel_tol=1e-9)
    assert ok and its == its_g
    assert (rn / r0) ** (1.0 / its) == pytest.approx(red_g, abs=5e-5)


@pytest.mark.parametrize("nq,quad", [(5, "gauss"), (4, "gauss"), (4, "gll")])
def test_operator_matches_kronecker_and_c_oracle(oracle_mod, nq, quad):
    fe = oracle_mod.fe
    p = 3
    mesh, dofs, rd, bas, G, JxW = _bp3_setup(fe, 5, p, nq, quad)  # 4x4x2 cells
    rng = np.random.default_rng(3)
    u = rng.standard_normal(rd["n_owned"])
    u[rd["constrained"]] = 0.0
    y = fe.op_apply(u, rd, bas, G)
    lat, dims = dofs["lattice_of_global"], dofs["dims"]
    U = np.zeros(dims[::-1])
    U[lat[:, 2], lat[:, 1], lat[:, 0]] = u
    Y = np.zeros(dims[::-1])
    Y[lat[:, 2], lat[:, 1], lat[:, 0]] = y
    if quad == "gauss":  # the GLL-collocated operator under-integrates: no Kronecker identity with exact mass
        Yk = fe.kron_apply(mesh, bas, U[1:-1, 1:-1, 1:-1])
        assert np.abs(Y[1:-1, 1:-1, 1:-1] - Yk).max() <= 1e-12 * np.abs(Yk).max()
    for flags, kw in ((1, dict(laplace=True)), (2, dict(laplace=False, mass=True)), (3, dict(laplace=True, mass=True))):
        y0 = fe.op_apply(u, rd, bas, G, JxW, **kw)
        y1 = oracle_mod.port.op_apply(u, nm=p + 1, nq=nq, collocated=(quad == "gll"), flags=flags,
                                      shape_values=bas["B"].T.copy(), co_shape_gradients=bas["D"].T.copy(),
                                      G=G, JxW=JxW, dof_indices=rd["dof_indices"],
                                      colors=_colors(mesh, rd["cells"]), constrained=rd["constrained"])
        assert np.abs(y0 - y1).max() <= 1e-13 * np.abs(y0).max()


def test_chebyshev_preconditioner_algebra():
    """oracle.fe.chebyshev_preconditioner: the recurrence reproduces the Chebyshev polynomial exactly -- on a diagonal
    operator the residual polynomial 1 - lambda p_k(lambda) equals T_k((theta - lambda) / delta) / T_k(theta / delta) -- and
    the power iteration finds the largest eigenvalue of D^-1 A (times the safety factor 1.2)."""
    import numpy as np
    from numpy.polynomial import chebyshev as T
    import oracle
    fe = oracle.fe
    lam = np.linspace(0.05, 2.0, 64)
    apply = lambda v: lam * v
    inv_diag = np.ones_like(lam)
    degree, lmax, rng = 5, 2.0, 20.0
    M = fe.chebyshev_preconditioner(apply, inv_diag, degree, lmax, rng)
    z = M(np.ones_like(lam))                       # p_k(lambda) on every eigenvalue at once
    theta, delta = 0.5 * (lmax + lmax / rng), 0.5 * (lmax - lmax / rng)
    Tk = lambda x: T.chebval(x, [0] * degree + [1])
    assert np.allclose(1.0 - lam * z, Tk((theta - lam) / delta) / Tk(theta / delta), rtol=1e-10, atol=1e-12)
    est = fe.estimate_max_eigenvalue(apply, inv_diag, len(lam), 200)
    assert est == pytest.approx(1.2 * 2.0, rel=1e-3)


def test_fast_single_rank_numbering_equals_literal_first_touch():
    """oracle.fe.rank_data_single_fast (what bench.py's CPU arm builds its mesh with) == the literal first-touch simulation."""
    import numpy as np
    import oracle
    fe = oracle.fe
    for sub, nref, p in [((1, 1, 1), 1, 1), ((2, 1, 1), 1, 2), ((2, 2, 1), 1, 3), ((1, 1, 1), 2, 2), ((1, 2, 1), 1, 4)]:
        m = fe.BoxMesh(sub, nref)
        a = fe.rank_data(m, fe.distribute_dofs(m, p, 1), 0)
        b = fe.rank_data_single_fast(m, p)
        assert np.array_equal(a["dof_indices"], b["dof_indices"]) and np.array_equal(a["constrained"], b["constrained"])
        assert a["n_owned"] == b["n_owned"] and b["n_ghost"] == 0


def test_reference_gpu_kernel_shims_export_their_entry_points():
    """oracle/_ref/libref_gpu_*.so (the reference's own CUDA kernels, compiled in place): present when /root/reference was
    there at build time, and exporting the entry points oracle/ref_gpu.py binds (no GPU call here)."""
    import ctypes
    import os
    from oracle import ref_gpu
    if not ref_gpu.available():
        pytest.skip("oracle/_ref/libref_gpu_*.so not built in this checkout")
    here = os.path.dirname(os.path.abspath(ref_gpu.__file__))
    assert hasattr(ctypes.CDLL(os.path.join(here, "_ref", "libref_gpu_ceedbk.so")), "ref_gpu_ceedbk")
    assert hasattr(ctypes.CDLL(os.path.join(here, "_ref", "libref_gpu_sumfact.so")), "ref_gpu_sumfact_bk1")


def test_p_transfer_oracle_is_exact_for_coarse_polynomials_and_adjoint():
    """oracle.fe.p_transfer: the prolongation reproduces every polynomial of the coarse degree at the fine nodes (an
    implementation-independent property of the FE_Q embedding), and the restriction is its exact transpose."""
    import numpy as np
    import oracle
    fe = oracle.fe
    m = fe.BoxMesh((2, 1, 1), 1)
    for pf, pc in ((4, 2), (3, 1), (8, 4)):
        rf, rc = fe.rank_data_single_fast(m, pf, dirichlet=False), fe.rank_data_single_fast(m, pc, dirichlet=False)
        P, R = fe.p_transfer(rf["dof_indices"], rc["dof_indices"], pf, pc, rf["n_owned"], rc["n_owned"])

        def coords(rd, p):
            t, _ = fe.gll_01(p + 1)
            n = p + 1
            X = np.zeros((rd["n_owned"], 3))
            l = np.arange(n ** 3)
            a, b, c = l % n, (l // n) % n, l // (n * n)
            for ci in range(m.n_cells):
                x, y, z = m.cell_xyz[ci]
                X[rd["dof_indices"][ci]] = np.stack([(x + t[a]) * m.h[0], (y + t[b]) * m.h[1], (z + t[c]) * m.h[2]], 1)
            return X
        f = lambda X: 1 + X[:, 0] ** pc * X[:, 1] - 0.5 * X[:, 2] ** pc * X[:, 0] + X[:, 1] * X[:, 2]
        assert np.abs(P(f(coords(rc, pc))) - f(coords(rf, pf))).max() <= 1e-12
        rng = np.random.default_rng(pf)
        u, r = rng.standard_normal(rc["n_owned"]), rng.standard_normal(rf["n_owned"])
        assert abs(P(u) @ r - u @ R(r)) <= 1e-11 * np.abs(P(u)).dot(np.abs(r))


@pytest.mark.parametrize("p,dq,quad", [(1, 2, "gauss"), (3, 2, "gauss"), (4, 1, "gauss"), (6, 1, "gll"), (8, 2, "gauss"), (5, 1, "gll")])
def test_separable_form_of_the_operator_on_axis_aligned_cells(oracle_mod, p, dq, quad):
    """The identity behind the separable ("cartesian") kernels, checked on the CPU against the oracle's cell kernel
    B^T D^T G D B (+ B^T JxW B): on an axis-aligned box cell the cell matrix is
        c_rr K x M x M + c_ss M x K x M + c_tt M x M x K  (+ det J M x M x M),   K = (D B)^T W (D B),  M = B^T W B,
    with c = det J diag(1/hz^2, 1/hy^2, 1/hx^2) (r <-> slowest local index <-> z); K and M are symmetric and point-symmetric
    (what the packed even-odd halves of csrc/eo_contract.h rely on); diagonal and int phi_i follow from diag K, diag M and
    m = B^T w.  Same formulas as b200fe_op_create / sumfact_cart.cuh / cart_diagonal_kernel / cart_rhs_kernel."""
    fe = oracle_mod.fe
    nq = p + dq
    bas = fe.basis_1d(p, nq, quad)
    B, D, w = bas["B"], bas["D"], bas["wq"]
    om = fe.BoxMesh((2, 1, 1), 0, p1=(0.0, -1.0, 0.5), p2=(0.6, -0.3, 0.9))   # two cells 0.3 x 0.7 x 0.4
    rd = fe.rank_data(om, fe.distribute_dofs(om, p, 1), 0)
    G, JxW = fe.geometric_factors(fe.cell_nodes(om, rd["cells"], 1), 1, bas)
    hx, hy, hz = 0.3, 0.7, 0.4
    det = hx * hy * hz
    c_rr, c_ss, c_tt = det / hz ** 2, det / hy ** 2, det / hx ** 2
    DB = D @ B
    K, M, m = DB.T @ (w[:, None] * DB), B.T @ (w[:, None] * B), B.T @ w
    for A in (K, M):
        assert np.abs(A - A.T).max() <= 1e-13 * np.abs(A).max()
        assert np.abs(A - A[::-1, ::-1]).max() <= 1e-12 * np.abs(A).max()
    nm = p + 1
    u = np.random.default_rng(p).standard_normal((len(rd["cells"]), nm, nm, nm))
    lap = (c_rr * np.einsum("ai,bj,ck,eijk->eabc", K, M, M, u) + c_ss * np.einsum("ai,bj,ck,eijk->eabc", M, K, M, u)
           + c_tt * np.einsum("ai,bj,ck,eijk->eabc", M, M, K, u))
    mass = det * np.einsum("ai,bj,ck,eijk->eabc", M, M, M, u)
    for kw, want in ((dict(laplace=True), lap), (dict(laplace=False, mass=True), mass), (dict(laplace=True, mass=True), lap + mass)):
        ref = fe.cell_kernel(u, bas, G, JxW, **kw)
        assert np.abs(ref - want).max() <= 1e-12 * np.abs(ref).max()
    # diagonal and rhs of the assembled operator
    dk, dm = np.diag(K), np.diag(M)
    diag_cell = (c_rr * np.einsum("a,b,c->abc", dk, dm, dm) + c_ss * np.einsum("a,b,c->abc", dm, dk, dm) + c_tt * np.einsum("a,b,c->abc", dm, dm, dk)
                 + det * np.einsum("a,b,c->abc", dm, dm, dm)).ravel()
    idx = rd["dof_indices"]
    valid = idx != 0xFFFFFFFF
    n_local = rd["n_owned"] + rd["n_ghost"]
    diag = np.bincount(idx[valid].astype(np.int64), weights=np.tile(diag_cell, (idx.shape[0], 1))[valid], minlength=n_local)
    diag[rd["constrained"]] = 1.0
    ref_diag = fe.op_diagonal(rd, bas, G, JxW, laplace=True, mass=True)
    assert np.abs(diag - ref_diag).max() <= 1e-11 * np.abs(ref_diag).max()
    rhs_cell = det * np.einsum("a,b,c->abc", m, m, m).ravel()
    rhs = np.bincount(idx[valid].astype(np.int64), weights=np.tile(rhs_cell, (idx.shape[0], 1))[valid], minlength=n_local)
    ref_rhs = fe.rhs_one(rd, bas, JxW)
    assert np.abs(rhs - ref_rhs).max() <= 1e-12 * np.abs(ref_rhs).max()
