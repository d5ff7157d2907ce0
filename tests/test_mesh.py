"""Host logic (no GPU): the product's closed-form box-mesh numbering / partition / ghost lists and
its 1-D bases against the oracle's literal simulation of deal.II (oracle/fe_oracle.py).
Bar: integer data bit-exact; 1-D matrices to 1e-13."""
import numpy as np
import pytest

import benchmarks_b200 as b


@pytest.mark.parametrize("p", range(1, 9))
def test_basis_matches_oracle(oracle_mod, p):
    fe = oracle_mod.fe
    for nq, quad, kind in ((p + 2, "gauss", b.QUAD_GAUSS), (p + 1, "gauss", b.QUAD_GAUSS), (p + 1, "gll", b.QUAD_GLL)):
        o = fe.basis_1d(p, nq, quad)
        m = b.basis_1d(p, nq, kind)
        assert np.abs(m["points"] - o["xq"]).max() < 1e-15
        assert np.abs(m["weights"] - o["wq"]).max() < 1e-15
        assert np.abs(m["shape_values"].reshape(p + 1, nq) - o["B"].T).max() < 1e-13
        assert np.abs(m["co_shape_gradients"].reshape(nq, nq) - o["D"].T).max() < 1e-11 * np.abs(o["D"]).max()
        assert np.abs(m["shape_gradients"].reshape(p + 1, nq) - o["Bg"].T).max() < 1e-11 * np.abs(o["Bg"]).max()


CASES = [  # subdivisions, n_refine, p, n_ranks, partition
    ((1, 1, 1), 0, 1, 1, "p4est"), ((1, 1, 1), 1, 2, 1, "p4est"), ((2, 2, 1), 1, 3, 1, "p4est"),
    ((2, 1, 1), 2, 2, 1, "p4est"), ((1, 1, 1), 2, 4, 2, "p4est"), ((2, 2, 1), 1, 2, 3, "p4est"),
    ((1, 1, 1), 2, 1, 8, "p4est"), ((2, 2, 2), 1, 3, 8, "p4est"), ((2, 1, 1), 1, 5, 4, "p4est"),
    ((1, 1, 5), 1, 2, 4, "blocks"), ((2, 2, 7), 0, 3, 3, "blocks"), ((1, 1, 1), 1, 8, 2, "p4est"),
]


@pytest.mark.parametrize("sub,nref,p,nranks,scheme", CASES)
@pytest.mark.parametrize("ghosts", ["minimal", "relevant"])
def test_numbering_partition_ghosts_bitexact(oracle_mod, sub, nref, p, nranks, scheme, ghosts):
    fe = oracle_mod.fe
    omesh = fe.BoxMesh(sub, nref)
    odofs = fe.distribute_dofs(omesh, p, nranks, scheme)
    for rank in range(nranks):
        ord_ = fe.rank_data(omesh, odofs, rank, ghost_set=ghosts)
        m = b.BoxMesh(sub, nref, p, n_ranks=nranks, rank=rank,
                      partition=b.PARTITION_P4EST if scheme == "p4est" else b.PARTITION_BLOCKS,
                      ghosts=b.GHOSTS_MINIMAL if ghosts == "minimal" else b.GHOSTS_RELEVANT)
        assert m.n_cells_global == omesh.n_cells
        assert m.n_dofs_global == len(odofs["lattice_of_global"])
        assert (m.owned_begin, m.owned_begin + m.n_owned) == odofs["owned_range"][rank]
        assert m.first_cell == (ord_["cells"][0] if len(ord_["cells"]) else m.first_cell)
        assert np.array_equal(m.cell_xyz, omesh.cell_xyz[ord_["cells"]])
        assert np.array_equal(m.ghost_global.astype(np.int64), ord_["ghost_global"])
        assert np.array_equal(m.ghost_owner, ord_["ghost_owner"])
        assert np.array_equal(m.dof_indices, ord_["dof_indices"])
        assert np.array_equal(m.constrained, ord_["constrained"])
        assert [int(x) for x in m.rank_dof_begin] == [r[0] for r in odofs["owned_range"]] + [odofs["owned_range"][-1][1]]


def test_bp3_sweep_dof_counts():
    """Global DoF counts of the reference's p=4 sweep (CEED_bp/results/1xGH200_P4.txt:636-642)."""
    expect = {8: (256, 18513), 9: (512, 35937), 10: (1024, 70785), 11: (2048, 139425), 12: (4096, 274625), 14: (16384, 1081665)}
    for cycle, (cells, dofs) in expect.items():
        m = b.BoxMesh.bp3_cycle(cycle, 4)
        assert (m.n_cells_global, m.n_dofs_global) == (cells, dofs)
        assert m.n_ghost == 0 and m.n_owned == dofs


def test_mesh_argument_errors():
    with pytest.raises(b.B200feError):
        b.BoxMesh((1, 1, 1), 0, 9)
    with pytest.raises(b.B200feError):
        b.BoxMesh((1, 1, 1), 0, 2, n_ranks=2, rank=0)  # fewer cells than ranks
    with pytest.raises(b.B200feError):
        b.BoxMesh((1, 0, 1), 0, 2)


@pytest.mark.parametrize("n_ranks,sub", [(2, (2, 1, 1)), (4, (2, 2, 1)), (8, (2, 2, 2))])
def test_bench_partitions_are_consistent(n_ranks, sub):
    """The weak-scaling layouts bench.py uses (one coarse cell of 64^3 per GPU; here 8^3): owned ranges tile the global
    numbering, every ghost lies in its owner's range, every rank owns a cube, and the overlap split is sane."""
    p, nref = 3, 3
    meshes = [b.BoxMesh(sub, nref, p, n_ranks=n_ranks, rank=r) for r in range(n_ranks)]
    n_global = meshes[0].n_dofs_global
    assert sum(m.n_owned for m in meshes) == n_global
    begins = [int(x) for x in meshes[0].rank_dof_begin]
    assert begins[0] == 0 and begins[-1] == n_global
    for r, m in enumerate(meshes):
        assert m.n_cells == 8 ** nref and (m.owned_begin, m.owned_begin + m.n_owned) == (begins[r], begins[r + 1])
        # one coarse cell per rank: the owned cells form a cube
        lo, hi = m.cell_xyz.min(axis=0), m.cell_xyz.max(axis=0)
        assert ((hi - lo + 1) == 8).all()
        own = m.ghost_owner
        assert (own != r).all()
        gg = m.ghost_global.astype(np.int64)
        assert ((gg >= np.array(begins)[own]) & (gg < np.array(begins)[own + 1])).all()
        assert (np.diff(gg) > 0).all()
        if n_ranks > 1:
            perm, n0, n1 = b.overlap_permutation(m.dof_indices, m.n_owned)
            assert n1 < m.n_cells and n0 + n1 <= m.n_cells
            # lower ranks own the interfaces: rank 0 never ghosts anything
            assert (r > 0) or m.n_ghost == 0


@pytest.mark.parametrize("sub,nref,p,nranks", [((1, 1, 1), 2, 2, 4), ((2, 2, 1), 1, 3, 3), ((2, 2, 2), 1, 2, 8), ((1, 1, 1), 1, 5, 2)])
def test_exchange_lists_without_communication_match_oracle(oracle_mod, sub, nref, p, nranks):
    """b200fe_exchange_* (every rank replays the other ranks' views; no message round) against the oracle's
    Partitioner-style lists: recv slices of the ghost segment, send lists in the peer's ghost order."""
    from benchmarks_b200.dist import exchange_lists_local
    fe = oracle_mod.fe
    om = fe.BoxMesh(sub, nref)
    od = fe.distribute_dofs(om, p, nranks)
    exs = fe.exchange_lists([fe.rank_data(om, od, r) for r in range(nranks)])
    for r in range(nranks):
        L = exchange_lists_local(b.BoxMesh(sub, nref, p, n_ranks=nranks, rank=r))
        ex = exs[r]
        assert sorted(ex["recv"]) == [int(t) for t, c in zip(L["peers"], L["recv_count"]) if c]
        assert sorted(ex["send"]) == [int(t) for t, c in zip(L["peers"], L["send_count"]) if c]
        for t, off, cnt, so, sc in zip(L["peers"], L["recv_offset"], L["recv_count"], L["send_offset"], L["send_count"]):
            if cnt:
                assert ex["recv"][int(t)] == (int(off), int(off + cnt))
            if sc:
                assert np.array_equal(ex["send"][int(t)], L["send_indices"][so:so + sc])


@pytest.mark.parametrize("sub,nref,p,lo,hi,nranks", [((2, 1, 1), 1, 3, (1, 0, 0), (3, 1, 2), 3), ((1, 1, 1), 2, 2, (1, 1, 1), (3, 3, 3), 4),
                                                     ((2, 1, 1), 1, 4, (0, 0, 0), (2, 2, 1), 8)])
def test_exchange_lists_without_communication_hanging(sub, nref, p, lo, hi, nranks):
    """Same for two-level meshes (the ghost sets carry the parents of hanging DoFs): against the lists derived from all
    ranks' mesh objects."""
    from benchmarks_b200.dist import exchange_lists_local
    ms = [b.HangingBoxMesh(sub, nref, p, lo, hi, n_ranks=nranks, rank=r) for r in range(nranks)]
    for r, m in enumerate(ms):
        L = exchange_lists_local(m)
        expect_peers = set(m.ghost_owner.tolist()) | {t for t in range(nranks) if t != r and (ms[t].ghost_owner == r).any()}
        assert [int(t) for t in L["peers"]] == sorted(expect_peers)
        for t, off, cnt, so, sc in zip(L["peers"], L["recv_offset"], L["recv_count"], L["send_offset"], L["send_count"]):
            t = int(t)
            sel = np.nonzero(m.ghost_owner == t)[0]
            assert cnt == len(sel) and (cnt == 0 or (sel[0] == off and sel[-1] == off + cnt - 1))
            want = (ms[t].ghost_global[ms[t].ghost_owner == r] - np.uint64(m.owned_begin)).astype(np.uint32)
            assert np.array_equal(want, L["send_indices"][so:so + sc])
