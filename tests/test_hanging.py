"""Hanging-node constraint handling on CPU (no GPU): the oracle's two-level FE_Q space (oracle/hanging_oracle.py,
conventions H1-H5) passes the patch tests that pin every constraint weight and index; the product's host-side
mesh builder (csrc/hangmesh.cc via b200fe_hangmesh_*) is bit-identical to the oracle's literal simulation for
numbering, index tables, ghost lists and constraint rows; and the multi-rank data reproduces the single-rank
operator.  Bar: integer data bit-exact; weights 1e-13; operator results 1e-12 relative."""
import numpy as np
import pytest

import benchmarks_b200 as b


def _setup(ho, fe, p, sub, nref, box, nranks=1, dirichlet=True, nq=None, p1=(-1.0, -1.0, -1.0), p2=None):
    mesh = ho.TwoLevelMesh(sub, nref, box, p1=p1, p2=p2)
    sp = ho.build_space(mesh, p, nranks, dirichlet=dirichlet)
    bas = fe.basis_1d(p, nq or p + 2, "gauss")
    return mesh, sp, bas


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5])
def test_patch_tests_pin_the_constraint_rows(oracle_mod, p):
    """The constrained space contains Q_p: interpolating a polynomial at the unconstrained nodes and distributing
    reproduces it at every hanging node (any wrong weight / parent index breaks this); the operator annihilates
    constants and returns the exact energy of a linear function."""
    fe, ho = oracle_mod.fe, oracle_mod.hanging
    mesh, sp, bas = _setup(ho, fe, p, (2, 2, 2), 0, ((0, 0, 0), (1, 1, 1)), dirichlet=False, p1=(0, 0, 0), p2=(2.0, 2.0, 2.0))
    rd = ho.rank_data(mesh, sp, 0)
    x, y, z = sp["pos"].T
    exact = (1 + x) ** p * (2 - y) ** min(p, 2) + z ** p * x - 3 * y * z
    u = exact.copy()
    u[sp["hanging"]] = 0.0
    assert len(sp["hanging"]) > 0
    assert np.abs(ho.distribute(rd, u) - exact).max() <= 1e-12 * np.abs(exact).max()
    G, JxW = fe.geometric_factors(ho.cell_nodes(mesh, rd["cells"], 1), 1, bas)
    free = np.ones(sp["n_dofs"], bool)
    free[rd["constrained"]] = False
    one = np.where(free, 1.0, 0.0)
    assert np.abs(ho.op_apply(rd, bas, G, one, JxW)[free]).max() <= 1e-13
    lin = np.where(free, x + 2 * y - z, 0.0)
    assert lin @ ho.op_apply(rd, bas, G, lin, JxW) == pytest.approx(6.0 * 8.0, rel=1e-13)  # |grad|^2 * volume
    # symmetry of C^T A C
    rng = np.random.default_rng(p)
    a, c = (np.where(free, rng.standard_normal(sp["n_dofs"]), 0.0) for _ in range(2))
    assert a @ ho.op_apply(rd, bas, G, c, JxW) == pytest.approx(c @ ho.op_apply(rd, bas, G, a, JxW), rel=1e-11)
    # rhs: int 1 over the domain = sum of b over a partition of unity restricted to free rows = volume
    assert ho.rhs_one(rd, bas, JxW).sum() == pytest.approx(8.0, rel=1e-13)


def test_dof_counts_follow_the_object_rule(oracle_mod):
    """H1: refining the corner cell of a 2x2x2 mesh adds (2p+1)^3 - 8 fine-level DoFs (everything of the children
    except the coarse cell's 8 vertices) and removes the coarse DoFs only that cell had (interior, 3 boundary faces,
    3 boundary edges)."""
    fe, ho = oracle_mod.fe, oracle_mod.hanging
    for p in (1, 2, 3):
        mesh, sp, _ = _setup(ho, fe, p, (2, 2, 2), 0, ((0, 0, 0), (1, 1, 1)))
        m = p - 1
        assert sp["n_dofs"] == (2 * p + 1) ** 3 - (m ** 3 + 3 * m ** 2 + 3 * m) + (2 * p + 1) ** 3 - 8
        # hanging: the fine nodes of the three inner faces off the domain boundary (union of three (2p)^2 grids),
        # minus the one coarse vertex among them
        assert len(sp["hanging"]) == 3 * (2 * p) ** 2 - 3 * (2 * p) + 1 - 1


CASES = [  # subdivisions, n_refine, p, refine_lo, refine_hi, n_ranks
    ((2, 2, 2), 0, 1, (0, 0, 0), (1, 1, 1), 1), ((2, 2, 2), 0, 2, (0, 0, 0), (1, 1, 1), 1),
    ((2, 2, 2), 0, 3, (1, 0, 1), (2, 1, 2), 1), ((1, 1, 1), 1, 2, (0, 0, 0), (1, 1, 1), 2),
    ((2, 1, 1), 1, 3, (1, 0, 0), (3, 1, 2), 3), ((1, 1, 1), 2, 2, (1, 1, 1), (3, 3, 3), 4),
    ((2, 1, 1), 1, 4, (0, 0, 0), (2, 2, 1), 8), ((1, 1, 1), 1, 5, (1, 1, 1), (2, 2, 2), 2),
    ((1, 1, 1), 1, 8, (0, 1, 0), (1, 2, 1), 1), ((3, 1, 1), 0, 6, (1, 0, 0), (2, 1, 1), 2),
    ((1, 1, 1), 1, 7, (0, 0, 0), (2, 2, 2), 3),  # everything refined: no hanging node at all
]


def _mesh_dict(m):
    return dict(owned_begin=int(m.owned_begin), n_owned=m.n_owned, n_ghost=m.n_ghost, ghost_global=m.ghost_global.astype(np.int64),
                dof_indices=m.dof_indices, constrained=m.constrained, hang_dof=m.hang_dof, hang_row_ptr=m.hang_row_ptr,
                hang_col=m.hang_col, hang_w=m.hang_w, cells=[tuple(int(v) for v in c) for c in m.cell_lxyz])


@pytest.mark.parametrize("sub,nref,p,lo,hi,nranks", CASES)
def test_product_mesh_bitexact_vs_oracle(oracle_mod, sub, nref, p, lo, hi, nranks):
    fe, ho = oracle_mod.fe, oracle_mod.hanging
    om = ho.TwoLevelMesh(sub, nref, (lo, hi))
    sp = ho.build_space(om, p, nranks)
    for r in range(nranks):
        rd = ho.rank_data(om, sp, r)
        m = b.HangingBoxMesh(sub, nref, p, lo, hi, n_ranks=nranks, rank=r)
        assert (m.n_cells_global, m.n_dofs_global) == (om.n_cells, sp["n_dofs"])
        assert (m.owned_begin, m.n_owned, m.n_ghost) == (rd["owned_begin"], rd["n_owned"], rd["n_ghost"])
        assert [int(v) for v in m.rank_dof_begin] == [a for a, _ in sp["owned_range"]] + [sp["n_dofs"]]
        assert np.array_equal(m.cell_lxyz, np.array(rd["cells"], dtype=np.int32).reshape(-1, 4))
        assert np.array_equal(m.dof_indices, rd["dof_indices"])
        assert np.array_equal(m.ghost_global.astype(np.int64), rd["ghost_global"])
        assert np.array_equal(m.ghost_owner, rd["ghost_owner"])
        assert np.array_equal(m.constrained, rd["constrained"])
        assert np.array_equal(m.hang_dof, rd["hang_dof"])
        assert np.array_equal(m.hang_row_ptr, rd["hang_row_ptr"])
        assert np.array_equal(m.hang_col, rd["hang_col"])
        assert np.abs(m.hang_w - rd["hang_w"]).max(initial=0.0) <= 1e-13


@pytest.mark.parametrize("sub,nref,p,lo,hi,nranks", [c for c in CASES if c[5] > 1 and c[2] <= 5])
def test_multirank_lists_reproduce_single_rank_operator(oracle_mod, sub, nref, p, lo, hi, nranks):
    """The product's per-rank tables (ghosts include the parents of hanging DoFs, rows in local indices) run through
    the vmult sequence give the single-rank result, compared through the numbering-independent DoF identities."""
    fe, ho = oracle_mod.fe, oracle_mod.hanging
    deform = lambda P: P + 0.03 * np.sin(2.0 * P[..., [1, 2, 0]])
    om = ho.TwoLevelMesh(sub, nref, (lo, hi))
    bas = fe.basis_1d(p, p + 2, "gauss")
    spP, sp1 = ho.build_space(om, p, nranks), ho.build_space(om, p, 1)
    ser_of = {k: d for d, k in enumerate(sp1["keys"])}
    perm = np.array([ser_of[k] for k in spP["keys"]])  # distributed global index -> serial global index
    rng = np.random.default_rng(11)
    u1 = rng.standard_normal(sp1["n_dofs"])
    rd1 = ho.rank_data(om, sp1, 0)
    G1, J1 = fe.geometric_factors(ho.cell_nodes(om, rd1["cells"], 2, deform), 2, bas)
    y1 = ho.op_apply(rd1, bas, G1, u1, J1, laplace=True, mass=True)
    rds = [_mesh_dict(b.HangingBoxMesh(sub, nref, p, lo, hi, n_ranks=nranks, rank=r)) for r in range(nranks)]
    geo = [fe.geometric_factors(ho.cell_nodes(om, rd["cells"], 2, deform), 2, bas) for rd in rds]
    yP = ho.distributed_apply(rds, bas, [g[0] for g in geo], u1[perm], [g[1] for g in geo], laplace=True, mass=True)
    assert np.abs(yP - y1[perm]).max() <= 1e-12 * np.abs(y1).max()


@pytest.mark.parametrize("sub,nref,p,lo,hi,nranks", CASES)
def test_face_blocks_bitexact_and_equivalent_to_rows(oracle_mod, sub, nref, p, lo, hi, nranks):
    """The face-structured form of the constraints (one block per coarse face, tensor-product trace interpolation):
    the product's blocks are bit-identical to the oracle's literal construction, every hanging row is the child of
    exactly one block, and distribute / condense through the blocks equal the CSR rows to rounding."""
    fe, ho = oracle_mod.fe, oracle_mod.hanging
    om = ho.TwoLevelMesh(sub, nref, (lo, hi))
    sp = ho.build_space(om, p, nranks)
    for r in range(nranks):
        rd = ho.rank_data(om, sp, r)
        par, chi = ho.face_blocks(om, sp, rd)
        m = b.HangingBoxMesh(sub, nref, p, lo, hi, n_ranks=nranks, rank=r)
        assert np.array_equal(m.face_parents, par) and np.array_equal(m.face_children, chi)
        assert np.abs(m.trace_weights - ho.trace_weights(p)).max() <= 1e-14
        kids = m.face_children[m.face_children != 0xFFFFFFFF]
        assert sorted(kids.tolist()) == sorted(m.hang_dof.tolist())
        u = np.random.default_rng(r).standard_normal(m.n_owned + m.n_ghost)
        md = _mesh_dict(m)
        assert np.abs(ho.distribute_faces(p, m.face_parents, m.face_children, u) - ho.distribute(md, u)).max(initial=0.0) <= 1e-13
        assert np.abs(ho.condense_faces(p, m.face_parents, m.face_children, u) - ho.condense(md, u)).max(initial=0.0) <= 1e-12


def test_bad_descriptors_are_rejected():
    with pytest.raises(b.B200feError):
        b.HangingBoxMesh((1, 1, 1), 1, 2, (0, 0, 0), (3, 1, 1))  # box outside the mesh
    with pytest.raises(b.B200feError):
        b.HangingBoxMesh((1, 1, 1), 1, 2, (1, 0, 0), (1, 1, 1))  # empty box
    with pytest.raises(b.B200feError):
        b.HangingBoxMesh((1, 1, 1), 1, 9, (0, 0, 0), (1, 1, 1))  # degree


def test_random_two_level_meshes_bitexact_vs_oracle(oracle_mod):
    """Property test over random small configurations (boxes, refinement regions, degrees, rank counts): every table of the
    product's mesh builder equals the oracle's literal simulation; face blocks cover every hanging DoF exactly once."""
    from hypothesis import given, settings, strategies as st
    fe, ho = oracle_mod.fe, oracle_mod.hanging

    @st.composite
    def configs(draw):
        sub = tuple(draw(st.integers(1, 2)) for _ in range(3))
        nref = draw(st.integers(0, 1 if sub[0] * sub[1] * sub[2] <= 2 else 0))
        cells = [s << nref for s in sub]
        lo = [draw(st.integers(0, c - 1)) for c in cells]
        hi = [draw(st.integers(l + 1, c)) for l, c in zip(lo, cells)]
        p = draw(st.integers(1, 3))
        n_cells = cells[0] * cells[1] * cells[2] + 7 * (hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2])
        nranks = draw(st.integers(1, min(5, n_cells)))
        return sub, nref, p, tuple(lo), tuple(hi), nranks

    @settings(max_examples=20, deadline=None, derandomize=True)
    @given(configs())
    def check(cfg):
        sub, nref, p, lo, hi, nranks = cfg
        om = ho.TwoLevelMesh(sub, nref, (lo, hi))
        sp = ho.build_space(om, p, nranks)
        for r in range(nranks):
            rd = ho.rank_data(om, sp, r)
            m = b.HangingBoxMesh(sub, nref, p, lo, hi, n_ranks=nranks, rank=r)
            assert (m.n_cells_global, m.n_dofs_global, m.owned_begin) == (om.n_cells, sp["n_dofs"], rd["owned_begin"])
            for name in ("dof_indices", "ghost_owner", "constrained", "hang_dof", "hang_row_ptr", "hang_col"):
                assert np.array_equal(getattr(m, name), rd[name]), (cfg, r, name)
            assert np.array_equal(m.ghost_global.astype(np.int64), rd["ghost_global"])
            assert np.abs(m.hang_w - rd["hang_w"]).max(initial=0.0) <= 1e-13
            par, chi = ho.face_blocks(om, sp, rd)
            assert np.array_equal(m.face_parents, par) and np.array_equal(m.face_children, chi), (cfg, r)

    check()


def test_golden_fixture_freezes_the_conventions(oracle_mod, golden_dir):
    """tests/golden/hanging_meshes.json (written by make_hanging_golden.py from the oracle): the oracle still produces it,
    and the product's builder matches the committed digests -- without importing the oracle for the product half."""
    import hashlib
    import json
    import os
    import sys
    sys.path.insert(0, golden_dir)
    import make_hanging_golden as mk
    gold = json.load(open(os.path.join(golden_dir, "hanging_meshes.json")))["cases"]
    dg = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    for want, case in zip(gold, mk.CASES):
        assert json.loads(json.dumps(mk.tables(case))) == want      # oracle half
        sub, nref, p, lo, hi, nranks = case
        for r, wr in enumerate(want["ranks"]):                     # product half
            m = b.HangingBoxMesh(sub, nref, p, lo, hi, n_ranks=nranks, rank=r)
            assert (m.n_cells_global, m.n_dofs_global) == (want["n_cells"], want["n_dofs"])
            assert (m.n_owned, m.n_ghost, int(m.owned_begin), m.n_cells) == (wr["n_owned"], wr["n_ghost"], wr["owned_begin"], wr["n_cells"])
            got = dict(dof_indices=dg(m.dof_indices), ghost_global=dg(m.ghost_global), constrained=dg(m.constrained), hang_dof=dg(m.hang_dof),
                       hang_row_ptr=dg(m.hang_row_ptr), hang_col=dg(m.hang_col), hang_w_rounded=dg(np.round(m.hang_w, 12) + 0.0),
                       face_parents=dg(m.face_parents), face_children=dg(m.face_children))
            for k, v in got.items():
                assert v == wr[k], (case, r, k)


def test_face_blocks_without_dirichlet_constraints(oracle_mod):
    """No Dirichlet DoFs (pure Neumann mesh): hanging DoFs on the domain boundary stay hanging and their boundary parents stay in
    the rows; the face form still covers every hanging DoF once and equals the rows."""
    ho = oracle_mod.hanging
    for sub, nref, p, lo, hi, nranks in (((2, 2, 2), 0, 2, (0, 0, 0), (1, 1, 1), 1), ((2, 1, 1), 1, 3, (1, 0, 0), (3, 1, 2), 3)):
        for r in range(nranks):
            m = b.HangingBoxMesh(sub, nref, p, lo, hi, n_ranks=nranks, rank=r, dirichlet=False)
            assert len(m.constrained) == len(m.hang_dof[m.hang_dof < m.n_owned])
            kids = m.face_children[m.face_children != 0xFFFFFFFF]
            assert sorted(kids.tolist()) == sorted(m.hang_dof.tolist())
            u = np.random.default_rng(r).standard_normal(m.n_owned + m.n_ghost)
            md = _mesh_dict(m)
            assert np.abs(ho.distribute_faces(p, m.face_parents, m.face_children, u) - ho.distribute(md, u)).max(initial=0.0) <= 1e-13
            assert np.abs(ho.condense_faces(p, m.face_parents, m.face_children, u) - ho.condense(md, u)).max(initial=0.0) <= 1e-12
