"""GPU parity of the L-vector operator path (gather -> sum factorisation -> atomic scatter, constrained
rows, diagonal, right-hand side) and of CG, through the C ABI, against the CPU oracle.

Tolerances (north star): one FP64 operator application <= 1e-12 relative max-norm; CG iteration
counts within +-1 of the oracle and, for p=4 on the reference's box meshes, equal +-1 to the goldens of
CEED_bp/results/1xGH200_P4.txt.  Atomics make the summation order nondeterministic, hence never bitwise.
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-12

VARIANTS = [  # (name, nq offset from p, quad, kind)
    ("bp3", 2, "gauss", "laplace"), ("bp35", 1, "gauss", "laplace"), ("bp5", 1, "gll", "laplace"),
    ("bp1", 2, "gauss", "mass"), ("helmholtz", 1, "gauss", "helmholtz"),
]
DEFORM = (0.05, 2.0)


def _oracle_setup(fe, sub, nref, p, nq, quad, p_geo, deform):
    om = fe.BoxMesh(sub, nref)
    od = fe.distribute_dofs(om, p, 1)
    rd = fe.rank_data(om, od, 0)
    bas = fe.basis_1d(p, nq, quad)
    dfm = None if deform is None else (lambda P: P + deform[0] * np.sin(deform[1] * P[..., [1, 2, 0]]))
    G, JxW = fe.geometric_factors(fe.cell_nodes(om, rd["cells"], p_geo, dfm), p_geo, bas)
    return om, od, rd, bas, G, JxW


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("p", range(1, 9))
@pytest.mark.parametrize("name,dq,quad,kind", VARIANTS)
def test_vmult_matches_oracle_on_deformed_mesh(oracle_mod, p, name, dq, quad, kind):
    import benchmarks_b200 as b
    fe = oracle_mod.fe
    sub, nref = ((2, 1, 1), 1) if p <= 5 else ((2, 1, 1), 0)
    nq = p + dq
    p_geo = min(p, 2)
    om, od, rd, bas, G, JxW = _oracle_setup(fe, sub, nref, p, nq, quad, p_geo, DEFORM)
    mesh = b.BoxMesh(sub, nref, p)
    A = b.LaplaceOperator(mesh, nq=nq, quad=quad, kind=kind, p_geo=p_geo, deform=DEFORM)
    # geometry (compute_G_tensors) parity
    if A.G is not None:
        assert rel(A.G.cpu().numpy().reshape(G.shape), G) <= TOL
    assert rel(A.JxW.cpu().numpy().reshape(JxW.shape), JxW) <= TOL
    rng = np.random.default_rng(p)
    src = rng.standard_normal(mesh.n_owned)
    ref = fe.op_apply(src, rd, bas, G, JxW, laplace=kind != "mass", mass=kind != "laplace")
    dst = torch.full((mesh.n_owned,), 7.0, dtype=torch.float64, device="cuda")  # vmult must overwrite dst
    A.vmult(dst, torch.from_numpy(src).cuda())
    assert rel(dst.cpu().numpy(), ref) <= TOL, name
    # fused inner product
    dst2 = A.initialize_dof_vector()
    dot = A.vmult_dot(dst2, torch.from_numpy(src).cuda())
    assert abs(dot.item() - float(src @ ref)) <= 1e-11 * np.abs(src).dot(np.abs(ref))


@pytest.mark.parametrize("name,dq,quad,kind", [VARIANTS[0], VARIANTS[2], VARIANTS[3]])
@pytest.mark.parametrize("p", [2, 3, 6])
def test_exclusive_interior_is_verified_on_the_index_table(oracle_mod, p, name, dq, quad, kind):
    """Plain stores for cell-interior DoFs (b200fe_op_exclusive_interior): on from p = 3 on an FE_Q table; an index table
    that maps the interior positions of two cells to the same DoFs must keep the atomic scatter -- and stay correct."""
    import benchmarks_b200 as b
    fe = oracle_mod.fe
    sub, nref, nq = (2, 1, 1), 1, p + dq
    om, od, rd, bas, G, JxW = _oracle_setup(fe, sub, nref, p, nq, quad, 1, DEFORM)
    mesh = b.BoxMesh(sub, nref, p)
    A = b.LaplaceOperator(mesh, nq=nq, quad=quad, kind=kind, deform=DEFORM)
    assert A.launch_info()["exclusive_interior"] == (1 if p >= 3 else 0)
    # cell 1 now writes to the DoFs of cell 0: every interior DoF of cell 0 has two writers
    mesh.dof_indices[1] = mesh.dof_indices[0]
    idx = rd["dof_indices"].copy()
    idx[1] = idx[0]
    assert np.array_equal(idx, mesh.dof_indices)
    A2 = b.LaplaceOperator(mesh, nq=nq, quad=quad, kind=kind, deform=DEFORM)
    assert A2.launch_info()["exclusive_interior"] == 0
    src = np.random.default_rng(10 + p).standard_normal(mesh.n_owned)
    ref = fe.op_apply(src, idx, bas, G, JxW, laplace=kind != "mass", mass=kind != "laplace", constrained=rd["constrained"])
    dst = torch.full((mesh.n_owned,), -3.0, dtype=torch.float64, device="cuda")
    A2.vmult(dst, torch.from_numpy(src).cuda())
    assert rel(dst.cpu().numpy(), ref) <= TOL
    # and the verified operator agrees with the oracle on the untouched table (plain stores on, dst pre-filled with garbage)
    ref1 = fe.op_apply(src, rd, bas, G, JxW, laplace=kind != "mass", mass=kind != "laplace")
    A.vmult(dst, torch.from_numpy(src).cuda())
    assert rel(dst.cpu().numpy(), ref1) <= TOL


@pytest.mark.parametrize("p,sub,nref,n_ranks,rank,dirichlet", [
    (1, (2, 1, 1), 1, 1, 0, True), (3, (1, 2, 1), 2, 3, 1, True), (6, (2, 2, 1), 1, 4, 3, True), (8, (1, 1, 1), 1, 2, 0, False),
    (4, (2, 2, 2), 2, 8, 5, True)])
def test_index_table_expanded_on_the_device_equals_host_table(p, sub, nref, n_ranks, rank, dirichlet):
    """b200fe_boxmesh_dof_indices_device (27 entity bases per cell expanded by a kernel) == b200fe_boxmesh_fill, bit for bit;
    the host table is in turn bit-identical to the oracle's first-touch simulation (tests/test_mesh.py)."""
    import benchmarks_b200 as b
    mesh = b.BoxMesh(sub, nref, p, n_ranks=n_ranks, rank=rank, dirichlet=dirichlet)
    assert mesh._dof_indices is None                       # nothing expanded on the host yet
    dev = mesh.dof_indices_device(torch.device("cuda", 0))
    assert mesh._dof_indices is None
    host = mesh.dof_indices
    assert dev.shape == host.shape and np.array_equal(dev.cpu().numpy(), host)


@pytest.mark.parametrize("p", [1, 2, 4, 7])
def test_diagonal_rhs_and_dummy(oracle_mod, p):
    import benchmarks_b200 as b
    fe = oracle_mod.fe
    sub, nref = (1, 1, 2), 1
    for name, dq, quad, kind in (VARIANTS[0], VARIANTS[2], VARIANTS[4]):
        nq = p + dq
        om, od, rd, bas, G, JxW = _oracle_setup(fe, sub, nref, p, nq, quad, 1, None)
        mesh = b.BoxMesh(sub, nref, p)
        A = b.LaplaceOperator(mesh, nq=nq, quad=quad, kind=kind)
        diag = A.compute_diagonal().cpu().numpy()
        ref = fe.op_diagonal(rd, bas, G, JxW, laplace=kind != "mass", mass=kind != "laplace")
        assert rel(diag, ref) <= 1e-11
        assert rel(A.compute_rhs().cpu().numpy(), fe.rhs_one(rd, bas, JxW)) <= TOL
    # vmult_dummy(ghost_exchange_on, computation_on) of portable_laplace_operator.h:175-235
    src = torch.randn(mesh.n_owned, dtype=torch.float64, device="cuda")
    full, comp_only, ghost_only = (torch.full_like(src, 3.0) for _ in range(3))
    A.vmult(full, src)
    A.vmult_dummy(comp_only, src, False, True)
    A.vmult_dummy(ghost_only, src, True, False)
    con = torch.from_numpy(mesh.constrained.astype(np.int64)).cuda()
    free = torch.ones_like(src, dtype=torch.bool)
    free[con] = False
    assert torch.equal(comp_only[free], full[free]) or (comp_only[free] - full[free]).abs().max() <= 1e-12 * full.abs().max()
    assert (comp_only[con] == 0).all()                  # computation only: constrained rows are not copied
    assert torch.equal(ghost_only[con], src[con]) and (ghost_only[free] == 3.0).all()  # no dst = 0 without computation


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("name,dq,quad", [("bp3", 2, "gauss"), ("bp35", 1, "gauss"), ("bp5", 1, "gll")])
def test_trilinear_on_the_fly_geometry_matches_oracle(oracle_mod, p, name, dq, quad):
    """geometry='trilinear' (SURVEY 8f.1): general hexahedra given by their 8 vertices, the Jacobian rebuilt at every
    quadrature point inside the kernel instead of streaming 48 nq^3 bytes of G per cell.  Against the oracle's operator with
    G from the same MappingQ1 cells (vertices displaced by the smooth map), <= 1e-12; same CG iteration count as stored G."""
    import benchmarks_b200 as b
    fe = oracle_mod.fe
    sub, nref, nq = (2, 1, 1), 1, p + dq
    om, od, rd, bas, G, JxW = _oracle_setup(fe, sub, nref, p, nq, quad, 1, DEFORM)
    mesh = b.BoxMesh(sub, nref, p)
    A = b.LaplaceOperator(mesh, nq=nq, quad=quad, p_geo=1, deform=DEFORM, geometry="trilinear", with_jxw=True)
    rng = np.random.default_rng(300 + p)
    src = rng.standard_normal(mesh.n_owned)
    ref = fe.op_apply(src, rd, bas, G)
    dst = A.initialize_dof_vector()
    dot = A.vmult_dot(dst, torch.from_numpy(src).cuda())
    assert rel(dst.cpu().numpy(), ref) <= TOL, name
    assert abs(dot.item() - float(src @ ref)) <= 1e-11 * np.abs(src).dot(np.abs(ref))
    A2 = b.LaplaceOperator(mesh, nq=nq, quad=quad, p_geo=1, deform=DEFORM)
    its = []
    for op in (A, A2):
        ctl = b.ReductionControl(5000, 1e-16, 1e-9)
        b.SolverCG(ctl).solve(op, op.initialize_dof_vector(), A2.compute_rhs())
        its.append(ctl.last_step())
    assert abs(its[0] - its[1]) <= 1


@pytest.mark.parametrize("p,name,dq,quad,kind", [(2, "bp3", 2, "gauss", "laplace"), (4, "bp5", 1, "gll", "laplace"), (3, "helmholtz", 1, "gauss", "helmholtz")])
def test_chebyshev_preconditioned_cg_matches_oracle(oracle_mod, p, name, dq, quad, kind):
    """PreconditionChebyshev (degree 4 polynomial in D^-1 A) inside SolverCG: eigenvalue estimate and iteration count
    against the numpy restatement of the same algorithm; fewer iterations than Jacobi."""
    import benchmarks_b200 as b
    fe = oracle_mod.fe
    sub, nref, nq = (2, 1, 1), 1, p + dq
    om, od, rd, bas, G, JxW = _oracle_setup(fe, sub, nref, p, nq, quad, 2, DEFORM)
    mesh = b.BoxMesh(sub, nref, p)
    A = b.LaplaceOperator(mesh, nq=nq, quad=quad, kind=kind, p_geo=2, deform=DEFORM)
    kw = dict(laplace=kind != "mass", mass=kind != "laplace")
    apply = lambda v: fe.op_apply(v, rd, bas, G, JxW, **kw)
    inv_diag = 1.0 / fe.op_diagonal(rd, bas, G, JxW, **kw)
    lam_ref = fe.estimate_max_eigenvalue(apply, inv_diag, mesh.n_owned, 12)
    P = b.PreconditionChebyshev(A, degree=4, smoothing_range=15.0, eig_iterations=12)
    assert P.max_eigenvalue == pytest.approx(lam_ref, rel=1e-9)
    rhs_ref = fe.rhs_one(rd, bas, JxW)
    M = fe.chebyshev_preconditioner(apply, inv_diag, 4, lam_ref, 15.0)
    _, its_ref, _, _, ok = fe.solver_cg(apply, rhs_ref, 500, 1e-16, 1e-9, precond=M)
    rhs, x = A.compute_rhs(), A.initialize_dof_vector()
    ctl = b.ReductionControl(500, 1e-16, 1e-9)
    b.SolverCG(ctl).solve(A, x, rhs, P)
    assert ok and abs(ctl.last_step() - its_ref) <= 1
    r = A.initialize_dof_vector()
    A.vmult(r, x)
    assert (rhs - r).norm().item() <= 5e-9 * rhs.norm().item()
    ctl_j = b.ReductionControl(2000, 1e-16, 1e-9)
    b.SolverCG(ctl_j).solve(A, A.initialize_dof_vector(), rhs, A.get_matrix_diagonal_inverse())
    assert ctl.last_step() < ctl_j.last_step()


@pytest.mark.parametrize("pf,pc", [(2, 1), (3, 1), (4, 2), (6, 3), (8, 4), (8, 7)])
def test_p_transfer_matches_oracle(oracle_mod, pf, pc):
    """MGTransferGlobalCoarsening for polynomial coarsening (CEED_bp/src/bp3.cc:27 includes its header): prolongate_and_add and
    restrict_and_add against the numpy restatement, Dirichlet DoFs masked on both levels; restriction is the exact transpose."""
    import benchmarks_b200 as b
    fe = oracle_mod.fe
    sub, nref = (2, 1, 1), 1
    om = fe.BoxMesh(sub, nref)
    rdf, rdc = fe.rank_data_single_fast(om, pf), fe.rank_data_single_fast(om, pc)
    P, R = fe.p_transfer(rdf["dof_indices"], rdc["dof_indices"], pf, pc, rdf["n_owned"], rdc["n_owned"])
    mf, mc = b.BoxMesh(sub, nref, pf), b.BoxMesh(sub, nref, pc)
    Af, Ac = b.LaplaceOperator(mf, quad="gll"), b.LaplaceOperator(mc, quad="gll")
    T = b.PTransfer(Af, Ac)
    rng = np.random.default_rng(pf * 10 + pc)
    uc, rf = rng.standard_normal(mc.n_owned), rng.standard_normal(mf.n_owned)
    uc[mc.constrained] = 0.0
    fine = torch.zeros(mf.n_owned, dtype=torch.float64, device="cuda")
    T.prolongate_and_add(fine, torch.from_numpy(uc).cuda())
    assert rel(fine.cpu().numpy(), P(uc)) <= TOL
    coarse = torch.zeros(mc.n_owned, dtype=torch.float64, device="cuda")
    T.restrict_and_add(coarse, torch.from_numpy(rf).cuda())
    assert rel(coarse.cpu().numpy(), R(rf)) <= TOL
    assert abs(float(fine.cpu().numpy() @ rf) - float(uc @ coarse.cpu().numpy())) <= 1e-11 * np.abs(fine.cpu().numpy()).dot(np.abs(rf))


def test_p_multigrid_preconditioned_cg_matches_oracle(oracle_mod):
    """V-cycle over the degrees 4, 2, 1 with Chebyshev smoothers as SolverCG preconditioner: iteration count against the numpy
    restatement of the same cycle (+-1), far fewer iterations than Jacobi, and a solved system."""
    import benchmarks_b200 as b
    fe = oracle_mod.fe
    sub, nref, degrees = (2, 1, 1), 1, (4, 2, 1)
    ops, levels, rds = [], [], []
    for p in degrees:
        om, od, rd, bas, G, JxW = _oracle_setup(fe, sub, nref, p, p + 1, "gll", 2, DEFORM)
        mesh = b.BoxMesh(sub, nref, p)
        A = b.LaplaceOperator(mesh, quad="gll", p_geo=2, deform=DEFORM)
        apply = (lambda rd_, bas_, G_: (lambda v: fe.op_apply(v, rd_, bas_, G_)))(rd, bas, G)
        inv_diag = 1.0 / fe.op_diagonal(rd, bas, G, JxW)
        levels.append((apply, inv_diag, fe.estimate_max_eigenvalue(apply, inv_diag, mesh.n_owned, 12)))
        ops.append(A)
        rds.append((rd, bas, JxW))
    transfers = [fe.p_transfer(rds[l][0]["dof_indices"], rds[l + 1][0]["dof_indices"], degrees[l], degrees[l + 1], rds[l][0]["n_owned"], rds[l + 1][0]["n_owned"])
                 for l in range(len(degrees) - 1)]
    M = fe.pmg_vcycle(levels, transfers, 3, 15.0, 8)
    rhs_ref = fe.rhs_one(*rds[0])
    _, its_ref, _, _, ok = fe.solver_cg(levels[0][0], rhs_ref, 200, 1e-16, 1e-9, precond=M)
    pre = b.PreconditionPMG(ops, smoother_degree=3, smoothing_range=15.0, coarse_degree=8, eig_iterations=12)
    assert pre.lambdas == pytest.approx([lv[2] for lv in levels], rel=1e-9)
    A = ops[0]
    # one V-cycle on a random residual
    r = np.random.default_rng(4).standard_normal(A.mesh.n_owned)
    r[A.mesh.constrained] = 0.0
    z = A.initialize_dof_vector()
    pre.vmult(z, torch.from_numpy(r).cuda())
    assert rel(z.cpu().numpy(), M(r)) <= 1e-10
    rhs, x = A.compute_rhs(), A.initialize_dof_vector()
    ctl = b.ReductionControl(200, 1e-16, 1e-9)
    b.SolverCG(ctl).solve(A, x, rhs, pre)
    assert ok and abs(ctl.last_step() - its_ref) <= 1
    res = A.initialize_dof_vector()
    A.vmult(res, x)
    assert (rhs - res).norm().item() <= 5e-9 * rhs.norm().item()
    ctl_j = b.ReductionControl(2000, 1e-16, 1e-9)
    b.SolverCG(ctl_j).solve(A, A.initialize_dof_vector(), rhs, A.get_matrix_diagonal_inverse())
    assert ctl.last_step() * 3 < ctl_j.last_step()


@pytest.fixture(scope="module")
def cg_golden(golden_dir):
    with open(os.path.join(golden_dir, "bp3_cg_p4.json")) as f:
        return json.load(f)["rows"]


@pytest.mark.parametrize("row", range(0, 13))
@pytest.mark.parametrize("nq", [6, 5])
def test_cg_reference_goldens_p4(cg_golden, row, nq):
    """bp3 protocol: rhs = int phi, x0 = 0, ReductionControl(1e9, 1e-16, 1e-9) (bp3.cc:266-285).  Rows 0-12 are every
    single-GPU row of CEED_bp/results/1xGH200_P4.txt:636-648 (18,513 ... 67,634,433 DoFs, 92 ... 1389 iterations); rows
    13-14 (135 M / 270 M DoFs, 4_GPU.out:779-780) are checked by bench.py's `parity` block on 8 GPUs."""
    import benchmarks_b200 as b
    cycle, cells, ndofs, its_g, red_g = cg_golden[row]
    if nq == 5 and row > 10:
        pytest.skip("the nq = 5 goldens are the same numbers; the two largest meshes run once (nq = 6)")
    mesh = b.BoxMesh.bp3_cycle(cycle, 4)
    assert (mesh.n_cells_global, mesh.n_dofs_global) == (cells, ndofs)
    A = b.LaplaceOperator(mesh, nq=nq)
    rhs = A.compute_rhs()
    x = A.initialize_dof_vector()
    ctl = b.ReductionControl(10 ** 9, 1e-16, 1e-9)
    b.SolverCG(ctl).solve(A, x, rhs)
    assert abs(ctl.last_step() - its_g) <= 1
    red = (ctl.last_value() / ctl.initial_value()) ** (1.0 / ctl.last_step())
    assert red == pytest.approx(red_g, abs=2e-3 if row < 7 else 5e-4)
    # the solve really solved: ||b - A x|| <= 1e-9 ||b|| (recomputed residual)
    r = A.initialize_dof_vector()
    A.vmult(r, x)
    assert (rhs - r).norm().item() <= 2e-9 * rhs.norm().item()


@pytest.mark.parametrize("p,name,dq,quad,kind,jacobi", [
    (2, "bp3", 2, "gauss", "laplace", False), (6, "bp5", 1, "gll", "laplace", False), (3, "helmholtz", 1, "gauss", "helmholtz", True),
    (5, "bp1", 2, "gauss", "mass", False), (8, "bp35", 1, "gauss", "laplace", True)])
def test_cg_matches_c_oracle(oracle_mod, p, name, dq, quad, kind, jacobi):
    import benchmarks_b200 as b
    fe = oracle_mod.fe
    sub, nref = ((2, 2, 1), 1) if p <= 6 else ((2, 1, 1), 1)
    nq = p + dq
    om, od, rd, bas, G, JxW = _oracle_setup(fe, sub, nref, p, nq, quad, 1, None)
    mesh = b.BoxMesh(sub, nref, p)
    A = b.LaplaceOperator(mesh, nq=nq, quad=quad, kind=kind)
    # bp5_kokkos right-hand side: b[i] = i % 8 on unconstrained owned DoFs (benchmark.cc:341-347)
    rhs = np.arange(mesh.n_owned, dtype=np.float64) % 8
    rhs[mesh.constrained] = 0.0
    flags = {"laplace": 1, "mass": 2, "helmholtz": 3}[kind]
    inv_diag = None
    pre = None
    if jacobi:
        pre = A.get_matrix_diagonal_inverse()
        inv_diag = pre.cpu().numpy()
    xo, its_o, r0_o, rn_o, ok_o = oracle_mod.port.cg_solve(
        rhs, nm=p + 1, nq=nq, collocated=(quad == "gll"), flags=flags, shape_values=bas["B"].T.copy(),
        co_shape_gradients=bas["D"].T.copy(), G=G, JxW=JxW, dof_indices=rd["dof_indices"], constrained=rd["constrained"],
        inv_diag=inv_diag, max_it=5000, abs_tol=1e-15, rel_tol=1e-8)
    x = A.initialize_dof_vector()
    ctl = b.ReductionControl(5000, 1e-15, 1e-8)
    b.SolverCG(ctl).solve(A, x, torch.from_numpy(rhs).cuda(), pre)
    assert ok_o and abs(ctl.last_step() - its_o) <= 1
    assert ctl.initial_value() == pytest.approx(r0_o, rel=1e-13)
    assert rel(x.cpu().numpy(), xo) <= 1e-6  # both stop at 1e-8 relative residual


def test_cg_no_convergence_is_reported_like_dealii():
    import benchmarks_b200 as b
    mesh = b.BoxMesh.bp3_cycle(8, 4)
    A = b.LaplaceOperator(mesh, nq=5)
    rhs = A.compute_rhs()
    x = A.initialize_dof_vector()
    ctl = b.ReductionControl(10, 1e-16, 1e-12)  # bp5_kokkos caps its at 100 and swallows NoConvergence
    with pytest.raises(b.NoConvergence):
        b.SolverCG(ctl).solve(A, x, rhs)
    assert ctl.last_step() == 10
    # host-buffer entry point gives the same iterate
    hx = torch.empty(mesh.n_owned, dtype=torch.float64).pin_memory()
    hb = rhs.cpu().pin_memory()
    with pytest.raises(b.NoConvergence):
        b.SolverCG(ctl).solve_host(A, hx, hb)
    assert (hx - x.cpu()).abs().max().item() <= 1e-12 * x.abs().max().item()


def _kron_full(fe, mesh, A, bas, u):
    """Element-free operator on a uniform box: (Kx x My x Mz + ...) u evaluated with torch matmuls."""
    import benchmarks_b200 as b
    p = mesh.p
    dev = u.device
    n = [c * p + 1 for c in mesh.cells]
    idx = A.dof_indices.view(mesh.n_cells, -1).to(torch.int64)
    xyz = torch.from_numpy(mesh.cell_xyz if A.perm is None else mesh.cell_xyz[A.perm]).to(dev).to(torch.int64)
    nm = p + 1
    l = torch.arange(nm ** 3, device=dev)
    a, bb, c = l % nm, (l // nm) % nm, l // (nm * nm)
    X = xyz[:, 0:1] * p + a[None]
    Y = xyz[:, 1:2] * p + bb[None]
    Z = xyz[:, 2:3] * p + c[None]
    valid = idx != 0xFFFFFFFF
    lin = (Z * n[1] + Y) * n[0] + X
    U = torch.zeros(n[2] * n[1] * n[0], dtype=torch.float64, device=dev)
    U[lin[valid]] = u[idx[valid]]
    U = U.view(n[2], n[1], n[0])[1:-1, 1:-1, 1:-1]
    mats = [[torch.from_numpy(m).to(dev) for m in fe.kron_1d(mesh.cells[d], mesh.h[d], bas)[:2]] for d in range(3)]
    (Kx, Mx), (Ky, My), (Kz, Mz) = mats
    t = lambda Az, Ay, Ax: torch.einsum("ai,ijk->ajk", Az, torch.einsum("bj,ijk->ibk", Ay, torch.einsum("ck,ijk->ijc", Ax, U)))
    Yk = t(Mz, My, Kx) + t(Mz, Ky, Mx) + t(Kz, My, Mx)
    return Yk, lin, valid, idx, n


@pytest.mark.parametrize("cells_log2,p,quad,dq", [(5, 4, "gauss", 2), (6, 6, "gll", 1)])
def test_full_size_apply_against_kronecker_form(oracle_mod, cells_log2, p, quad, dq):
    """BASELINE configs C4 per-GPU size (p=4 at 32^3 here, 64^3 in bench) and C3 (64^3 cells, p=6,
    57,066,625 DoFs): the whole vmult against the element-free Kronecker operator,
    an oracle that shares no gather/scatter or sum-factorisation code with the kernel."""
    import benchmarks_b200 as b
    fe = oracle_mod.fe
    mesh = b.BoxMesh((1, 1, 1), cells_log2, p)
    nq = p + dq
    A = b.LaplaceOperator(mesh, nq=nq, quad=quad, with_jxw=False)
    bas = fe.basis_1d(p, nq, quad)
    gen = torch.Generator(device="cuda").manual_seed(1)
    u = torch.randn(mesh.n_owned, dtype=torch.float64, device="cuda", generator=gen)
    u[torch.from_numpy(mesh.constrained.astype(np.int64)).cuda()] = 0.0
    y = A.initialize_dof_vector()
    A.vmult(y, u)
    Yk, lin, valid, idx, n = _kron_full(fe, mesh, A, bas, u)
    Yp = torch.zeros(n[2] * n[1] * n[0], dtype=torch.float64, device="cuda")
    Yp[lin[valid]] = y[idx[valid]]
    Yp = Yp.view(n[2], n[1], n[0])[1:-1, 1:-1, 1:-1]
    assert ((Yp - Yk).abs().max() / Yk.abs().max()).item() <= TOL
    # symmetry and linearity at full size
    v = torch.randn(mesh.n_owned, dtype=torch.float64, device="cuda", generator=gen)
    Av = A.initialize_dof_vector()
    A.vmult(Av, v)
    assert abs(torch.dot(y, v).item() - torch.dot(u, Av).item()) <= 1e-12 * (y.norm() * v.norm()).item()


@pytest.mark.parametrize("p,dq,quad,kind", [(3, 2, "gauss", "laplace"), (2, 1, "gll", "laplace"), (4, 1, "gauss", "helmholtz"), (6, 1, "gll", "laplace")])
def test_tail_batches_do_not_leak_stale_shared_memory(oracle_mod, p, dq, quad, kind):
    """Cell counts that are not a multiple of the CTA's element batch: the unused slots read a stale
    (here NaN-poisoned) G buffer and must not reach dst or the fused inner product (round-1 bug found
    by the 8-GPU run: 0 * NaN from a tail batch turned p.Ap into NaN)."""
    import benchmarks_b200 as b
    from benchmarks_b200._lib import check, lib
    fe = oracle_mod.fe
    sub, nref = (3, 1, 1), 0 if p >= 6 else 1   # 3 or 24 cells: never a multiple of the batch sizes in use
    nq = p + dq
    om, od, rd, bas, G, JxW = _oracle_setup(fe, sub, nref, p, nq, quad, 1, None)
    mesh = b.BoxMesh(sub, nref, p)
    A = b.LaplaceOperator(mesh, nq=nq, quad=quad, kind=kind)
    src = np.random.default_rng(11).standard_normal(mesh.n_owned)
    ref = fe.op_apply(src, rd, bas, G, JxW, laplace=kind != "mass", mass=kind != "laplace")
    for _ in range(3):
        check(lib.b200fe_debug_poison_smem(None))
        dst = A.initialize_dof_vector()
        dot = A.vmult_dot(dst, torch.from_numpy(src).cuda())
        assert torch.isfinite(dst).all() and np.isfinite(dot.item())
        assert rel(dst.cpu().numpy(), ref) <= TOL
        assert abs(dot.item() - float(src @ ref)) <= 1e-11 * np.abs(src).dot(np.abs(ref))


def test_cg_stops_on_nan_like_dealii():
    import benchmarks_b200 as b
    mesh = b.BoxMesh.bp3_cycle(6, 2)
    A = b.LaplaceOperator(mesh)
    rhs = A.compute_rhs()
    rhs[5] = float("nan")
    x = A.initialize_dof_vector()
    ctl = b.ReductionControl(10 ** 9, 1e-16, 1e-9)
    with pytest.raises(b.NoConvergence):
        b.SolverCG(ctl).solve(A, x, rhs)
    assert ctl.last_step() <= 1


def test_vector_valued_apply_bp6_style(oracle_mod):
    """CEED BP6-style vector Laplacian (3 components, GLL collocated): component-blocked apply equals the
    scalar oracle applied per component (no reference implementation exists for BP2/4/6)."""
    import benchmarks_b200 as b
    fe = oracle_mod.fe
    p, sub, nref = 4, (2, 1, 1), 1
    om, od, rd, bas, G, JxW = _oracle_setup(fe, sub, nref, p, p + 1, "gll", 2, DEFORM)
    mesh = b.BoxMesh(sub, nref, p)
    A = b.LaplaceOperator(mesh, quad="gll", p_geo=2, deform=DEFORM)
    rng = np.random.default_rng(2)
    src = rng.standard_normal((3, mesh.n_owned))
    dst = torch.empty(3 * mesh.n_owned, dtype=torch.float64, device="cuda")
    A.vmult_components(dst, torch.from_numpy(src.ravel()).cuda(), 3)
    out = dst.cpu().numpy().reshape(3, -1)
    for c in range(3):
        assert rel(out[c], fe.op_apply(src[c], rd, bas, G)) <= TOL


@pytest.mark.parametrize("p", [1, 2, 3, 4, 6, 8])
@pytest.mark.parametrize("dq,quad", [(2, "gauss"), (1, "gauss"), (1, "gll")])
def test_on_the_fly_affine_geometry_matches_stored_G_and_oracle(oracle_mod, p, dq, quad):
    """SURVEY section 8f.1 ("next"): geometric factors from six per-cell constants on affine (here anisotropic box)
    cells.  Same operator as the stored-G path and as the oracle, to 1e-12; same CG iteration count."""
    import benchmarks_b200 as b
    fe = oracle_mod.fe
    sub, nref = ((3, 2, 1), 1) if p <= 4 else ((3, 1, 1), 0)
    p1, p2 = (0.0, -1.0, 0.5), (1.3, 0.2, 1.0)   # cells 0.217 x 0.3 x 0.25: anisotropic, off-origin
    nq = p + dq
    om = fe.BoxMesh(sub, nref, p1=p1, p2=p2)
    od = fe.distribute_dofs(om, p, 1)
    rd = fe.rank_data(om, od, 0)
    bas = fe.basis_1d(p, nq, quad)
    G, JxW = fe.geometric_factors(fe.cell_nodes(om, rd["cells"], 1), 1, bas)
    mesh = b.BoxMesh(sub, nref, p, p1=p1, p2=p2)
    A_st = b.LaplaceOperator(mesh, nq=nq, quad=quad)
    A_af = b.LaplaceOperator(mesh, nq=nq, quad=quad, geometry="affine")
    # axis-aligned cells: the separable ("cartesian") kernels -- three contractions at the points for the collocated operator,
    # seven on the nodal values for the interpolated ones
    assert A_af.launch_info()["cartesian"] == 1
    src = np.random.default_rng(p).standard_normal(mesh.n_owned)
    ref = fe.op_apply(src, rd, bas, G)
    d_src = torch.from_numpy(src).cuda()
    y_st, y_af = A_st.initialize_dof_vector(), A_af.initialize_dof_vector()
    A_st.vmult(y_st, d_src)
    dot = A_af.vmult_dot(y_af, d_src)
    assert rel(y_af.cpu().numpy(), ref) <= TOL
    assert rel(y_af.cpu().numpy(), y_st.cpu().numpy()) <= TOL
    assert abs(dot.item() - float(src @ ref)) <= 1e-11 * np.abs(src).dot(np.abs(ref))
    rhs = A_af.compute_rhs()
    its = []
    for A in (A_st, A_af):
        x = A.initialize_dof_vector()
        ctl = b.ReductionControl(5000, 1e-16, 1e-9)
        b.SolverCG(ctl).solve(A, x, rhs)
        its.append(ctl.last_step())
    assert abs(its[0] - its[1]) <= 1


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("dq,kind", [(2, "mass"), (1, "helmholtz")])
def test_separable_mass_and_helmholtz_on_axis_aligned_cells(oracle_mod, p, dq, kind):
    """BP1 (mass, QGauss(p+2)) and bp5_kokkos' Helmholtz operator (QGauss(p+1)) with the geometry on the fly: on axis-aligned
    cells the mass term det J (M x M x M) is separable too.  Apply, diagonal (Jacobi) and rhs against the oracle and against
    the stored-geometry operator; sheared cells are refused for these operators."""
    import benchmarks_b200 as b
    fe = oracle_mod.fe
    sub, nref = ((3, 2, 1), 1) if p <= 4 else ((3, 1, 1), 0)
    p1, p2 = (0.0, -1.0, 0.5), (1.3, 0.2, 1.0)   # cells 0.217 x 0.3 x 0.25: anisotropic, off-origin
    nq = p + dq
    om = fe.BoxMesh(sub, nref, p1=p1, p2=p2)
    rd = fe.rank_data(om, fe.distribute_dofs(om, p, 1), 0)
    bas = fe.basis_1d(p, nq, "gauss")
    G, JxW = fe.geometric_factors(fe.cell_nodes(om, rd["cells"], 1), 1, bas)
    lap, mass = kind != "mass", True
    mesh = b.BoxMesh(sub, nref, p, p1=p1, p2=p2)
    A = b.LaplaceOperator(mesh, nq=nq, quad="gauss", kind=kind, geometry="affine", with_jxw=False)
    assert A.launch_info()["cartesian"] == 1 and A.JxW is None and A.G is None
    src = np.random.default_rng(70 + p).standard_normal(mesh.n_owned)
    ref = fe.op_apply(src, rd, bas, G, JxW, laplace=lap, mass=mass)
    y = A.initialize_dof_vector()
    dot = A.vmult_dot(y, torch.from_numpy(src).cuda())
    assert rel(y.cpu().numpy(), ref) <= TOL
    assert abs(dot.item() - float(src @ ref)) <= 1e-11 * np.abs(src).dot(np.abs(ref))
    assert rel(A.compute_diagonal().cpu().numpy(), fe.op_diagonal(rd, bas, G, JxW, laplace=lap, mass=mass)) <= 1e-11
    assert rel(A.compute_rhs().cpu().numpy(), fe.rhs_one(rd, bas, JxW)) <= TOL
    # Jacobi-preconditioned CG: same iteration count as the stored-geometry operator
    A_st = b.LaplaceOperator(mesh, nq=nq, quad="gauss", kind=kind)
    rhs = A_st.compute_rhs()
    its = []
    for op in (A_st, A):
        x = op.initialize_dof_vector()
        ctl = b.ReductionControl(5000, 1e-16, 1e-9)
        b.SolverCG(ctl).solve(op, x, rhs, op.get_matrix_diagonal_inverse())
        its.append(ctl.last_step())
    assert abs(its[0] - its[1]) <= 1
    shear_t = lambda N: torch.stack([N[:, 0] + 0.3 * N[:, 1], N[:, 1], N[:, 2]], dim=1)
    with pytest.raises(b.B200feError):
        b.LaplaceOperator(mesh, nq=nq, quad="gauss", kind=kind, geometry="affine", with_jxw=False, node_transform=shear_t)


@pytest.mark.parametrize("p", [1, 2, 3, 5, 6, 7, 8])
@pytest.mark.parametrize("dq,quad", [(2, "gauss"), (1, "gll")])
def test_on_the_fly_affine_geometry_on_sheared_cells(oracle_mod, monkeypatch, p, dq, quad):
    """Parallelepiped (sheared) cells: all six per-cell constants are non-zero, so the coupling terms of the general affine
    kernel are exercised and the cartesian test must say no; with B200FE_CARTESIAN=0 an axis-aligned mesh also stays on the
    general affine kernel and gives the cartesian kernel's result."""
    import benchmarks_b200 as b
    fe = oracle_mod.fe
    sub, nref = ((2, 1, 1), 1) if p <= 4 else ((2, 1, 1), 0)
    nq = p + dq
    shear_np = lambda P: np.stack([P[..., 0] + 0.3 * P[..., 1] - 0.2 * P[..., 2], P[..., 1] + 0.25 * P[..., 2], P[..., 2] + 0.1 * P[..., 0]], axis=-1)
    shear_t = lambda N: torch.stack([N[:, 0] + 0.3 * N[:, 1] - 0.2 * N[:, 2], N[:, 1] + 0.25 * N[:, 2], N[:, 2] + 0.1 * N[:, 0]], dim=1)
    om = fe.BoxMesh(sub, nref)
    rd = fe.rank_data(om, fe.distribute_dofs(om, p, 1), 0)
    bas = fe.basis_1d(p, nq, quad)
    G, _ = fe.geometric_factors(fe.cell_nodes(om, rd["cells"], 1, shear_np), 1, bas)
    mesh = b.BoxMesh(sub, nref, p)
    A = b.LaplaceOperator(mesh, nq=nq, quad=quad, geometry="affine", node_transform=shear_t)
    assert A.launch_info()["cartesian"] == 0
    src = np.random.default_rng(40 + p).standard_normal(mesh.n_owned)
    d_src = torch.from_numpy(src).cuda()
    y = A.initialize_dof_vector()
    A.vmult(y, d_src)
    assert rel(y.cpu().numpy(), fe.op_apply(src, rd, bas, G)) <= TOL
    if True:  # same axis-aligned mesh through the separable and the general affine kernel
        A_cart = b.LaplaceOperator(mesh, nq=nq, quad=quad, geometry="affine")
        monkeypatch.setenv("B200FE_CARTESIAN", "0")
        A_gen = b.LaplaceOperator(mesh, nq=nq, quad=quad, geometry="affine")
        monkeypatch.delenv("B200FE_CARTESIAN")
        assert A_cart.launch_info()["cartesian"] == 1 and A_gen.launch_info()["cartesian"] == 0
        y1, y2 = A_cart.initialize_dof_vector(), A_gen.initialize_dof_vector()
        d1, d2 = A_cart.vmult_dot(y1, d_src), A_gen.vmult_dot(y2, d_src)
        assert rel(y1.cpu().numpy(), y2.cpu().numpy()) <= TOL
        assert abs(d1.item() - d2.item()) <= 1e-11 * abs(d2.item())


@pytest.mark.parametrize("p,quad,dq", [(4, "gauss", 2), (6, "gll", 1)])
def test_cg_graph_replay_on_explicit_stream_equals_plain_launches(p, quad, dq):
    """On a non-default stream the library replays chunks of CG iterations as a CUDA graph; iterations queued after
    convergence are no-ops.  Same iteration count, same iterate as the plain-launch path (default stream)."""
    import benchmarks_b200 as b
    mesh = b.BoxMesh.bp3_cycle(9, p)
    A = b.LaplaceOperator(mesh, nq=p + dq, quad=quad)
    rhs = A.compute_rhs()
    x0, x1 = A.initialize_dof_vector(), A.initialize_dof_vector()
    c0, c1 = b.ReductionControl(10 ** 6, 1e-16, 1e-9), b.ReductionControl(10 ** 6, 1e-16, 1e-9)
    b.SolverCG(c0).solve(A, x0, rhs)                       # default stream: plain launches
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        b.SolverCG(c1).solve(A, x1, rhs, stream=side)      # explicit stream: graph replay
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    assert c1.last_step() == c0.last_step()
    assert (x1 - x0).abs().max().item() <= 1e-9 * x0.abs().max().item()
    # capped solve (bp5_kokkos style): the tail that does not fill a chunk runs un-graphed, the count is exact
    c2 = b.ReductionControl(37, 0.0, 0.0)
    with torch.cuda.stream(side):
        with pytest.raises(b.NoConvergence):
            b.SolverCG(c2).solve(A, x1, rhs, stream=side)
    assert c2.last_step() == 37
