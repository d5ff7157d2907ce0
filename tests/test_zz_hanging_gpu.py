"""GPU parity of the constrained operator C^T A C on two-level meshes with hanging nodes (BASELINE config C5,
north-star subsystem 1) and of the vector-valued CG (BP2/BP4/BP6), through the C ABI, against the CPU oracle
(oracle/hanging_oracle.py; "parity unpinned": the reference holds no hanging-node or vector-valued code).

Tolerances: one FP64 application <= 1e-12 relative max-norm; CG iteration counts within +-1 of the oracle's loop.
(The file sorts last on purpose: these paths are the newest.)
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-12
DEFORM = (0.04, 2.0)


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _setup(oracle_mod, p, sub, nref, lo, hi, nq, quad, kind, p_geo=2, deform=DEFORM, constraints="faces"):
    import benchmarks_b200 as b
    fe, ho = oracle_mod.fe, oracle_mod.hanging
    om = ho.TwoLevelMesh(sub, nref, (lo, hi))
    sp = ho.build_space(om, p, 1)
    rd = ho.rank_data(om, sp, 0)
    bas = fe.basis_1d(p, nq, quad)
    dfm = None if deform is None else (lambda P: P + deform[0] * np.sin(deform[1] * P[..., [1, 2, 0]]))
    G, JxW = fe.geometric_factors(ho.cell_nodes(om, rd["cells"], p_geo, dfm), p_geo, bas)
    mesh = b.HangingBoxMesh(sub, nref, p, lo, hi)
    A = b.LaplaceOperator(mesh, nq=nq, quad=quad, kind=kind, p_geo=p_geo, deform=deform, constraints=constraints)
    return fe, ho, rd, bas, G, JxW, mesh, A


VARIANTS = [("bp3", 2, "gauss", "laplace"), ("bp5", 1, "gll", "laplace"), ("bp1", 2, "gauss", "mass"), ("helmholtz", 1, "gauss", "helmholtz")]


@pytest.mark.parametrize("p", [1, 2, 3, 4, 6, 8])
@pytest.mark.parametrize("name,dq,quad,kind", VARIANTS)
@pytest.mark.parametrize("constraints", ["rows", "faces"])
def test_constrained_vmult_matches_oracle(oracle_mod, p, name, dq, quad, kind, constraints):
    """constraints = "rows": general CSR rows (AffineConstraints); "faces": tensor-product form per coarse face (default)."""
    sub, nref, lo, hi = ((2, 2, 1), 0, (0, 0, 0), (1, 2, 1)) if p >= 6 else ((1, 1, 1), 1, (1, 0, 1), (2, 1, 2))
    fe, ho, rd, bas, G, JxW, mesh, A = _setup(oracle_mod, p, sub, nref, lo, hi, p + dq, quad, kind, constraints=constraints)
    assert len(mesh.hang_dof) > 0
    assert rel(A.JxW.cpu().numpy().reshape(JxW.shape), JxW) <= TOL  # children are half-size cells of the same map
    rng = np.random.default_rng(100 + p)
    src = rng.standard_normal(mesh.n_owned)  # arbitrary values on the hanging entries too: rows act as identity
    ref = ho.op_apply(rd, bas, G, src, JxW, laplace=kind != "mass", mass=kind != "laplace")
    d_src = torch.from_numpy(src).cuda()
    dst = torch.full((mesh.n_owned,), 7.0, dtype=torch.float64, device="cuda")
    A.vmult(dst, d_src)
    assert rel(dst.cpu().numpy(), ref) <= TOL, name
    assert torch.equal(d_src.cpu(), torch.from_numpy(src))  # the hanging entries of src were restored
    dst2 = A.initialize_dof_vector()
    dot = A.vmult_dot(dst2, d_src)
    assert abs(dot.item() - float(src @ ref)) <= 1e-11 * np.abs(src).dot(np.abs(ref))
    # distribute (AffineConstraints::distribute) and the right-hand side b = C^T int phi
    x = d_src.clone()
    A.distribute(x)
    assert rel(x.cpu().numpy(), ho.distribute(rd, src)) <= TOL
    assert rel(A.compute_rhs().cpu().numpy(), ho.rhs_one(rd, bas, JxW)) <= TOL


@pytest.mark.parametrize("p,name,dq,quad,kind,constraints", [(1, "bp3", 2, "gauss", "laplace", "faces"), (2, "bp5", 1, "gll", "laplace", "rows"),
                                                             (3, "helmholtz", 1, "gauss", "helmholtz", "faces"), (2, "bp1", 2, "gauss", "mass", "faces"),
                                                             (4, "bp5", 1, "gll", "laplace", "faces")])
def test_diagonal_with_constraints_matches_oracle(oracle_mod, p, name, dq, quad, kind, constraints):
    """compute_diagonal of C^T A C (bp5_kokkos/benchmark.cc:218-251 semantics on a mesh with hanging nodes): every entry
    against the oracle's constrained operator applied to unit vectors; 1 on Dirichlet and hanging rows."""
    fe, ho, rd, bas, G, JxW, mesh, A = _setup(oracle_mod, p, (1, 1, 1), 1, (1, 0, 1), (2, 1, 2), p + dq, quad, kind, constraints=constraints)
    n = mesh.n_owned
    ref = np.empty(n)
    for j in range(n):
        e = np.zeros(n)
        e[j] = 1.0
        ref[j] = ho.op_apply(rd, bas, G, e, JxW, laplace=kind != "mass", mass=kind != "laplace")[j]
    diag = A.compute_diagonal().cpu().numpy()[:n]
    assert rel(diag, ref) <= TOL, name
    con = np.concatenate([mesh.constrained, mesh.hang_dof]).astype(np.int64)
    assert (diag[con] == 1.0).all()
    # and the Jacobi-preconditioned CG on the constrained operator runs with it
    import benchmarks_b200 as b
    rhs, x = A.compute_rhs(), A.initialize_dof_vector()
    ctl = b.ReductionControl(2000, 1e-16, 1e-9)
    b.SolverCG(ctl).solve(A, x, rhs, A.get_matrix_diagonal_inverse())
    assert ctl.last_step() > 0


@pytest.mark.parametrize("p,dq,quad,constraints", [(2, 2, "gauss", "faces"), (4, 1, "gll", "rows"), (7, 1, "gll", "faces")])
def test_cg_on_hanging_mesh_matches_oracle_loop(oracle_mod, p, dq, quad, constraints):
    """bp3 protocol (rhs = int phi, x0 = 0, ReductionControl(., 1e-16, 1e-9)) on the constrained operator; the solution is
    compared after distributing the constraints."""
    import benchmarks_b200 as b
    sub, nref, lo, hi = ((2, 2, 1), 0, (1, 0, 0), (2, 1, 1)) if p >= 6 else ((1, 1, 1), 1, (0, 0, 0), (1, 1, 2))
    fe, ho, rd, bas, G, JxW, mesh, A = _setup(oracle_mod, p, sub, nref, lo, hi, p + dq, quad, "laplace", constraints=constraints)
    rhs_o = ho.rhs_one(rd, bas, JxW)
    xo, its_o, r0_o, rn_o, ok_o = fe.solver_cg(lambda u: ho.op_apply(rd, bas, G, u, JxW), rhs_o, 5000, 1e-16, 1e-9)
    rhs = A.compute_rhs()
    x = A.initialize_dof_vector()
    ctl = b.ReductionControl(5000, 1e-16, 1e-9)
    b.SolverCG(ctl).solve(A, x, rhs)
    assert ok_o and abs(ctl.last_step() - its_o) <= 1
    assert ctl.initial_value() == pytest.approx(r0_o, rel=1e-12)
    A.distribute(x)
    assert rel(x.cpu().numpy(), ho.distribute(rd, xo)) <= 1e-6
    # explicit stream: CUDA-graph replay of iteration chunks with the constraint kernels inside
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    x1 = A.initialize_dof_vector()
    c1 = b.ReductionControl(5000, 1e-16, 1e-9)
    with torch.cuda.stream(side):
        b.SolverCG(c1).solve(A, x1, rhs, stream=side)
    torch.cuda.synchronize()
    assert abs(c1.last_step() - ctl.last_step()) <= 1


@pytest.mark.parametrize("hanging", [False, True])
def test_vector_valued_cg_bp6_style(oracle_mod, hanging):
    """One CG on the 3-component block system (shared alpha / beta), component-blocked vectors: iteration count and
    solution against the oracle's CG loop run on the stacked operator."""
    import benchmarks_b200 as b
    fe, ho = oracle_mod.fe, oracle_mod.hanging
    p, nc = 3, 3
    if hanging:
        _, _, rd, bas, G, JxW, mesh, A = _setup(oracle_mod, p, (1, 1, 1), 1, (0, 0, 0), (1, 1, 1), p + 1, "gll", "laplace")
        apply1 = lambda u: ho.op_apply(rd, bas, G, u, JxW)
    else:
        om = fe.BoxMesh((2, 1, 1), 1)
        od = fe.distribute_dofs(om, p, 1)
        rd = fe.rank_data(om, od, 0)
        bas = fe.basis_1d(p, p + 1, "gll")
        G, JxW = fe.geometric_factors(fe.cell_nodes(om, rd["cells"], 1), 1, bas)
        mesh = b.BoxMesh((2, 1, 1), 1, p)
        A = b.LaplaceOperator(mesh, quad="gll")
        apply1 = lambda u: fe.op_apply(u, rd, bas, G)
    n = mesh.n_owned
    rng = np.random.default_rng(5)
    rhs = rng.standard_normal((nc, n))
    rhs[:, mesh.constrained] = 0.0
    stacked = lambda u: np.concatenate([apply1(u[c * n:(c + 1) * n]) for c in range(nc)])
    xo, its_o, r0_o, _, ok_o = fe.solver_cg(stacked, rhs.ravel().copy(), 5000, 1e-16, 1e-9)
    x = torch.zeros(nc * n, dtype=torch.float64, device="cuda")
    ctl = b.ReductionControl(5000, 1e-16, 1e-9)
    b.SolverCG(ctl).solve(A, x, torch.from_numpy(rhs.ravel()).cuda(), n_components=nc)
    assert ok_o and abs(ctl.last_step() - its_o) <= 1
    assert ctl.initial_value() == pytest.approx(r0_o, rel=1e-12)
    assert rel(x.cpu().numpy(), xo) <= 1e-6
    # the component-blocked apply agrees with the per-component oracle on the same operator
    y = torch.empty_like(x)
    A.vmult_components(y, x, nc)
    assert rel(y.cpu().numpy(), stacked(x.cpu().numpy())) <= TOL


def test_constraint_rows_are_validated():
    import benchmarks_b200 as b
    from benchmarks_b200._lib import lib
    mesh = b.BoxMesh((1, 1, 1), 1, 2)
    A = b.LaplaceOperator(mesh)
    u32 = lambda *v: np.array(v, dtype=np.uint32)
    ptr = lambda a: a.ctypes.data
    w = np.array([0.5, 0.5])
    # chain: row 0 constrains DoF 5 to (6, 7), row 1 constrains 6
    hd, rp, col, w4 = u32(5, 6), u32(0, 2, 4), u32(6, 7, 8, 9), np.array([0.5, 0.5, 0.5, 0.5])
    assert lib.b200fe_op_set_constraints(A._h, 2, ptr(hd), ptr(rp), ptr(col), ptr(w4)) == 1
    hd, rp, col = u32(5), u32(0, 2), u32(6, mesh.n_owned + 3)  # parent outside the local vector
    assert lib.b200fe_op_set_constraints(A._h, 1, ptr(hd), ptr(rp), ptr(col), ptr(w)) == 1
    # a valid row is accepted, and the diagonal of C^T A C is available with the rows attached
    hd, rp, col = u32(13), u32(0, 2), u32(6, 7)
    assert lib.b200fe_op_set_constraints(A._h, 1, ptr(hd), ptr(rp), ptr(col), ptr(w)) == 0
    assert np.isfinite(A.compute_diagonal().cpu().numpy()).all()


def test_bp6_driver_matches_python_path():
    """C++ host layer (include/b200fe/operator.hpp: HangingBoxMesh, component-blocked Vector, SolverCG) through the bp6
    driver: mesh sizes and CG iteration counts equal the Python mirror's on the same meshes."""
    import os
    import subprocess
    import benchmarks_b200 as b
    drv = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "benchmarks_b200", "drivers")
    exe = os.path.join(drv, "bp6")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", drv], check=True)
    r = subprocess.run([exe, "2", "1", "12000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    rows = [[x.strip() for x in l.split("|")] for l in r.stdout.splitlines() if l.count("|") == 8 and l.split("|")[0].strip().isdigit()]
    assert len(rows) >= 2
    for row, cycle in zip(rows, (3, 4)):
        n_refine, rem = cycle // 3, cycle % 3
        sub = [2 if d < rem else 1 for d in range(3)]
        hi = [(s << n_refine) // 2 for s in sub]
        mesh = b.HangingBoxMesh(sub, n_refine, 2, (0, 0, 0), hi)
        assert (int(row[0]), int(row[1]), int(row[2])) == (mesh.n_cells_global, 3 * mesh.n_dofs_global, len(mesh.hang_dof))
        A = b.LaplaceOperator(mesh, quad="gll", p_geo=2, deform=(0.05, 2.0))
        rhs1 = A.compute_rhs()
        rhs = torch.cat([rhs1, rhs1, rhs1])
        x = torch.zeros_like(rhs)
        ctl = b.ReductionControl(10 ** 9, 1e-16, 1e-9)
        b.SolverCG(ctl).solve(A, x, rhs, n_components=3)
        assert abs(int(row[6]) - ctl.last_step()) <= 1


def test_geometry_from_dealii_inv_jacobian_views(oracle_mod):
    """G from MatrixFree's inv_jacobian(q, cell, ref, real) / JxW(q, cell) (Kokkos LayoutLeft), checked
    against the oracle's G of the same deformed cells, with K = J^-1 produced independently by numpy."""
    import ctypes as C
    import benchmarks_b200 as b
    from benchmarks_b200._lib import check, lib
    fe = oracle_mod.fe
    p, nq, p_geo = 3, 5, 2
    om = fe.BoxMesh((2, 1, 1), 1)
    cells = np.arange(om.n_cells)
    bas = fe.basis_1d(p, nq)
    nodes = fe.cell_nodes(om, cells, p_geo, lambda P: P + 0.05 * np.sin(2.0 * P[..., [1, 2, 0]]))
    G, JxW = fe.geometric_factors(nodes, p_geo, bas)
    # J at the points from the mapping nodes (same einsums as the oracle), K = J^-1 as deal.II stores it
    tg, _ = fe.gll_01(p_geo + 1)
    V, dV = fe.lagrange_values(tg, bas["xq"]), fe.lagrange_derivs(tg, bas["xq"])
    J = np.stack([np.einsum("cdzyx,rz,qy,px->cdrqp", nodes, V, V, dV), np.einsum("cdzyx,rz,qy,px->cdrqp", nodes, V, dV, V),
                  np.einsum("cdzyx,rz,qy,px->cdrqp", nodes, dV, V, V)], axis=2)          # [c, real, ref, z, y, x]
    K = np.linalg.inv(np.moveaxis(J, (1, 2), (-2, -1)))                                   # [c, z, y, x, ref, real]
    nc, nq3 = om.n_cells, nq ** 3
    view = np.ascontiguousarray(np.transpose(K.reshape(nc, nq3, 3, 3), (3, 2, 0, 1)))      # (real, ref, cell, q): q fastest
    d_K, d_J = torch.from_numpy(view.ravel()).cuda(), torch.from_numpy(JxW.ravel().copy()).cuda()
    d_G = torch.empty(nc * 6 * nq3, dtype=torch.float64, device="cuda")
    vp = lambda t: C.c_void_p(t.data_ptr())
    check(lib.b200fe_geometry_from_inv_jacobian(nc, nq, vp(d_K), vp(d_J), vp(d_G), None))
    assert rel(d_G.cpu().numpy().reshape(G.shape), G) <= TOL


@pytest.mark.parametrize("p", [1, 2, 4, 7, 8])
def test_face_structured_constraints_equal_rows(oracle_mod, p):
    """constraints="faces" (one CTA per coarse face, W (x) W in shared memory) against the oracle and the CSR-row path."""
    import benchmarks_b200 as b
    sub, nref, lo, hi = ((2, 2, 1), 0, (0, 0, 0), (1, 2, 1)) if p >= 6 else ((1, 1, 1), 1, (1, 0, 1), (2, 1, 2))
    fe, ho, rd, bas, G, JxW, mesh, A = _setup(oracle_mod, p, sub, nref, lo, hi, p + 1, "gll", "laplace", constraints="rows")
    Af = b.LaplaceOperator(mesh, nq=p + 1, quad="gll", p_geo=2, deform=DEFORM, constraints="faces")
    src = np.random.default_rng(p).standard_normal(3 * mesh.n_owned)
    ref = np.concatenate([ho.op_apply(rd, bas, G, src[c * mesh.n_owned:(c + 1) * mesh.n_owned], JxW) for c in range(3)])
    d_src = torch.from_numpy(src).cuda()
    y, yf = torch.empty_like(d_src), torch.empty_like(d_src)
    A.vmult_components(y, d_src, 3)
    Af.vmult_components(yf, d_src, 3)
    assert rel(yf.cpu().numpy(), ref) <= TOL and rel(yf.cpu().numpy(), y.cpu().numpy()) <= TOL
    assert torch.equal(d_src.cpu(), torch.from_numpy(src))
    assert rel(Af.compute_rhs().cpu().numpy(), ho.rhs_one(rd, bas, JxW)) <= TOL
    x = d_src[: mesh.n_owned].clone()
    Af.distribute(x)
    assert rel(x.cpu().numpy(), ho.distribute(rd, src[: mesh.n_owned])) <= TOL
    rhs, xs = Af.compute_rhs(), [Af.initialize_dof_vector(), A.initialize_dof_vector()]
    its = []
    for op, xv in zip((Af, A), xs):
        ctl = b.ReductionControl(5000, 1e-16, 1e-9)
        b.SolverCG(ctl).solve(op, xv, rhs)
        its.append(ctl.last_step())
    assert abs(its[0] - its[1]) <= 1


@pytest.mark.parametrize("p,dq,quad,kind", [(2, 2, "gauss", "laplace"), (4, 1, "gll", "laplace"), (3, 1, "gauss", "helmholtz"), (6, 1, "gll", "laplace")])
@pytest.mark.parametrize("constraints", ["rows", "faces"])
def test_separable_kernels_under_hanging_node_constraints(oracle_mod, p, dq, quad, kind, constraints):
    """Undeformed two-level mesh: parents and children are axis-aligned boxes of two sizes, so the geometry can be evaluated
    on the fly by the separable kernels (six constants per cell); C^T A C around them (distribute / condense, exclusive
    interior stores) must equal the oracle's constrained operator and the stored-geometry one."""
    import benchmarks_b200 as b
    sub, nref, lo, hi = ((2, 2, 1), 0, (0, 0, 0), (1, 2, 1)) if p >= 6 else ((1, 1, 1), 1, (1, 0, 1), (2, 1, 2))
    fe, ho, rd, bas, G, JxW, mesh, A_st = _setup(oracle_mod, p, sub, nref, lo, hi, p + dq, quad, kind, p_geo=1, deform=None, constraints=constraints)
    A = b.LaplaceOperator(mesh, nq=p + dq, quad=quad, kind=kind, p_geo=1, geometry="affine", with_jxw=False, constraints=constraints)
    assert A.launch_info()["cartesian"] == 1 and len(mesh.hang_dof) > 0
    src = np.random.default_rng(300 + p).standard_normal(mesh.n_owned)
    ref = ho.op_apply(rd, bas, G, src, JxW, laplace=kind != "mass", mass=kind != "laplace")
    d_src = torch.from_numpy(src).cuda()
    y, y_st = torch.full((mesh.n_owned,), -2.0, dtype=torch.float64, device="cuda"), A_st.initialize_dof_vector()
    A.vmult(y, d_src)
    A_st.vmult(y_st, d_src)
    assert rel(y.cpu().numpy(), ref) <= TOL
    assert rel(y.cpu().numpy(), y_st.cpu().numpy()) <= TOL
    assert rel(A.compute_rhs().cpu().numpy(), ho.rhs_one(rd, bas, JxW)) <= TOL
    # CG on the constrained operator: same iteration count with either geometry
    rhs = A_st.compute_rhs()
    its = []
    for op in (A_st, A):
        x = op.initialize_dof_vector()
        ctl = b.ReductionControl(5000, 1e-16, 1e-9)
        b.SolverCG(ctl).solve(op, x, rhs)
        its.append(ctl.last_step())
    assert abs(its[0] - its[1]) <= 1
