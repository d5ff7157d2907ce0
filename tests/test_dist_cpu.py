"""N>1 host logic on CPU: world_size-2 (and 3) gloo processes build their mesh partitions and the
exchange lists; checked against the oracle's single-process construction, and a gloo emulation of
update_ghost_values / compress(add) reproduces the single-rank operator result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, sub, nref, p, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import benchmarks_b200 as b
        from benchmarks_b200.dist import exchange_lists
        import oracle
        fe = oracle.fe
        mesh = b.BoxMesh(sub, nref, p, n_ranks=world, rank=rank)
        L = exchange_lists(mesh)
        # oracle view of the same partition
        om = fe.BoxMesh(sub, nref)
        od = fe.distribute_dofs(om, p, world)
        rds = [fe.rank_data(om, od, r) for r in range(world)]
        ex = fe.exchange_lists(rds)[rank]
        ok = sorted(ex["recv"]) == [t for t, c in zip(L["peers"], L["recv_count"]) if c]
        for t, off, cnt, soff, scnt in zip(L["peers"], L["recv_offset"], L["recv_count"], L["send_offset"], L["send_count"]):
            if cnt:
                ok &= ex["recv"][int(t)] == (int(off), int(off + cnt))
            if scnt:
                ok &= np.array_equal(ex["send"][int(t)], L["send_indices"][soff:soff + scnt])
            else:
                ok &= int(t) not in ex["send"]
        # distributed apply with gloo standing in for NCCL: update ghosts, local oracle apply, compress(add)
        bas = fe.basis_1d(p, p + 2)
        rd = rds[rank]
        G, _ = fe.geometric_factors(fe.cell_nodes(om, rd["cells"], 1), 1, bas)
        rng = np.random.default_rng(7)
        u_global = rng.standard_normal(len(od["lattice_of_global"]))
        v = np.zeros(mesh.n_owned + mesh.n_ghost)
        v[:mesh.n_owned] = u_global[mesh.owned_begin:mesh.owned_begin + mesh.n_owned]

        def exchange(send_bufs):
            out = {}
            reqs = []
            for t in L["peers"]:
                t = int(t)
                if t in send_bufs:
                    reqs.append(dist.isend(torch.from_numpy(send_bufs[t].copy()), t))
            for t, n in recv_sizes.items():
                out[t] = torch.empty(n, dtype=torch.float64)
                reqs.append(dist.irecv(out[t], t))
            for r in reqs:
                r.wait()
            return {t: o.numpy() for t, o in out.items()}

        # update_ghost_values
        send = {int(t): v[L["send_indices"][so:so + sc]] for t, so, sc in zip(L["peers"], L["send_offset"], L["send_count"]) if sc}
        recv_sizes = {int(t): int(c) for t, c in zip(L["peers"], L["recv_count"]) if c}
        got = exchange(send)
        for t, off, cnt in zip(L["peers"], L["recv_offset"], L["recv_count"]):
            if cnt:
                v[mesh.n_owned + off:mesh.n_owned + off + cnt] = got[int(t)]
        ok &= np.array_equal(v[mesh.n_owned:], u_global[mesh.ghost_global.astype(np.int64)])
        y = fe.op_apply(v, rd, bas, G, constrained=np.zeros(0, np.uint32))
        # compress(add): ghost contributions travel back
        send = {int(t): y[mesh.n_owned + off:mesh.n_owned + off + cnt] for t, off, cnt in zip(L["peers"], L["recv_offset"], L["recv_count"]) if cnt}
        recv_sizes = {int(t): int(c) for t, c in zip(L["peers"], L["send_count"]) if c}
        got = exchange(send)
        for t, so, sc in zip(L["peers"], L["send_offset"], L["send_count"]):
            if sc:
                np.add.at(y, L["send_indices"][so:so + sc], got[int(t)])
        y = y[:mesh.n_owned]
        y[mesh.constrained] = v[mesh.constrained]
        # single-rank reference
        od1 = fe.distribute_dofs(om, p, 1)
        rd1 = fe.rank_data(om, od1, 0)
        G1, _ = fe.geometric_factors(fe.cell_nodes(om, rd1["cells"], 1), 1, bas)
        # map the distributed numbering onto the serial one through lattice coordinates
        lat = od["lattice_of_global"]
        ser_of_dist = od1["global_of_lattice"][lat[:, 0], lat[:, 1], lat[:, 2]]
        u1 = np.empty_like(u_global)
        u1[ser_of_dist] = u_global
        y1 = fe.op_apply(u1, rd1, bas, G1)
        mine = np.arange(mesh.owned_begin, mesh.owned_begin + mesh.n_owned)
        err = np.abs(y - y1[ser_of_dist[mine]]).max() / np.abs(y1).max()
        q.put((rank, bool(ok), float(err)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,sub,nref,p", [(2, (1, 1, 1), 1, 2), (2, (2, 1, 1), 1, 3), (3, (2, 2, 1), 1, 2)])
def test_exchange_lists_and_distributed_apply_gloo(world, sub, nref, p):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, sub, nref, p, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
    assert len(results) == world
    for rank, ok, err in results:
        assert ok, f"rank {rank}: exchange lists differ from the oracle"
        assert err <= 1e-13, f"rank {rank}: distributed apply differs ({err})"


def test_overlap_permutation_splits_cells():
    import benchmarks_b200 as b
    mesh = b.BoxMesh((1, 1, 1), 2, 2, n_ranks=2, rank=1)
    perm, n0, n1 = b.overlap_permutation(mesh.dof_indices, mesh.n_owned)
    assert sorted(perm.tolist()) == list(range(mesh.n_cells))
    idx = mesh.dof_indices[perm]
    touches = ((idx >= mesh.n_owned) & (idx != 0xFFFFFFFF)).any(axis=1)
    assert not touches[:n0].any() and touches[n0:n0 + n1].all() and not touches[n0 + n1:].any()
    assert n1 > 0


@pytest.mark.parametrize("one_sided", [False, True])
def test_overlap_permutation_from_per_cell_flags_equals_the_table_path(one_sided):
    """The operator computes the "touches a ghost DoF" flags on the device from the device-side index table (signed view:
    INVALID = -1, valid indices < 2^31) and hands them to overlap_permutation: same permutation as from the host table."""
    import benchmarks_b200 as b
    for rank in range(3):
        mesh = b.BoxMesh((2, 1, 1), 2, 3, n_ranks=3, rank=rank)
        idx = mesh.dof_indices
        ref = b.overlap_permutation(idx, mesh.n_owned, one_sided=one_sided)
        signed = torch.from_numpy(idx).view(torch.int32)              # what LaplaceOperator does on the device tensor
        touches = (signed >= mesh.n_owned).any(dim=1).numpy()
        got = b.overlap_permutation(None, mesh.n_owned, one_sided=one_sided, touches=touches)
        assert np.array_equal(ref[0], got[0]) and ref[1:] == got[1:]


def _worker_hanging(rank, world, port, sub, nref, p, lo, hi, q):
    """Same emulation on a two-level mesh with hanging nodes: the ghost set must also carry the parents of the
    hanging DoFs, and distribute / condense run between the exchanges (the sequence of b200fe_op_vmult)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import benchmarks_b200 as b
        from benchmarks_b200.dist import exchange_lists
        import oracle
        fe, ho = oracle.fe, oracle.hanging
        mesh = b.HangingBoxMesh(sub, nref, p, lo, hi, n_ranks=world, rank=rank)
        L = exchange_lists(mesh)
        om = ho.TwoLevelMesh(sub, nref, (lo, hi))
        spP, sp1 = ho.build_space(om, p, world), ho.build_space(om, p, 1)
        ser_of = {k: d for d, k in enumerate(sp1["keys"])}
        perm = np.array([ser_of[k] for k in spP["keys"]])
        bas = fe.basis_1d(p, p + 2)
        rd = dict(dof_indices=mesh.dof_indices, hang_dof=mesh.hang_dof, hang_row_ptr=mesh.hang_row_ptr, hang_col=mesh.hang_col,
                  hang_w=mesh.hang_w)
        cells = [tuple(int(v) for v in c) for c in mesh.cell_lxyz]
        G, _ = fe.geometric_factors(ho.cell_nodes(om, cells, 1), 1, bas)
        u1 = np.random.default_rng(3).standard_normal(sp1["n_dofs"])
        u_global = u1[perm]
        n_own = mesh.n_owned
        v = np.zeros(n_own + mesh.n_ghost)
        v[:n_own] = u_global[mesh.owned_begin:mesh.owned_begin + n_own]

        def exchange(send_bufs, recv_sizes):
            out, reqs = {}, []
            for t, buf in send_bufs.items():
                reqs.append(dist.isend(torch.from_numpy(buf.copy()), t))
            for t, n in recv_sizes.items():
                out[t] = torch.empty(n, dtype=torch.float64)
                reqs.append(dist.irecv(out[t], t))
            for r in reqs:
                r.wait()
            return {t: o.numpy() for t, o in out.items()}

        peers = [int(t) for t in L["peers"]]
        send_sl = {t: L["send_indices"][so:so + sc] for t, so, sc in zip(peers, L["send_offset"], L["send_count"]) if sc}
        recv_sl = {t: (n_own + int(off), int(cnt)) for t, off, cnt in zip(peers, L["recv_offset"], L["recv_count"]) if cnt}
        got = exchange({t: v[ix] for t, ix in send_sl.items()}, {t: c for t, (_, c) in recv_sl.items()})  # update_ghost_values
        for t, (o, c) in recv_sl.items():
            v[o:o + c] = got[t]
        ok = bool(np.array_equal(v[n_own:], u_global[mesh.ghost_global.astype(np.int64)]))
        uh = ho.distribute(rd, v)
        y = ho.condense(rd, fe.op_apply(uh, mesh.dof_indices, bas, G, n_local=len(v), constrained=np.zeros(0, np.uint32)))
        got = exchange({t: y[o:o + c] for t, (o, c) in recv_sl.items()}, {t: len(ix) for t, ix in send_sl.items()})  # compress(add)
        for t, ix in send_sl.items():
            np.add.at(y, ix, got[t])
        y = y[:n_own]
        y[mesh.constrained] = v[mesh.constrained]
        rd1 = ho.rank_data(om, sp1, 0)
        G1, _ = fe.geometric_factors(ho.cell_nodes(om, rd1["cells"], 1), 1, bas)
        y1 = ho.op_apply(rd1, bas, G1, u1)
        mine = np.arange(mesh.owned_begin, mesh.owned_begin + n_own)
        err = np.abs(y - y1[perm[mine]]).max() / np.abs(y1).max()
        q.put((rank, ok, float(err)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,sub,nref,p,lo,hi", [(2, (1, 1, 1), 1, 2, (0, 0, 0), (1, 1, 1)), (3, (2, 1, 1), 1, 3, (1, 0, 0), (3, 1, 2))])
def test_hanging_mesh_exchange_and_distributed_apply_gloo(world, sub, nref, p, lo, hi):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_hanging, args=(r, world, port, sub, nref, p, lo, hi, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
    assert len(results) == world
    for rank, ok, err in results:
        assert ok, f"rank {rank}: ghost values differ after update_ghost_values"
        assert err <= 1e-12, f"rank {rank}: distributed constrained apply differs ({err})"


@pytest.mark.parametrize("world", [2, 3])
def test_cxx_communicator_and_exchange_lists(world):
    """C++ host layer (include/b200fe/operator.hpp) as `world` processes, no GPU: b200fe::Communicator hands every rank the
    same NCCL id (rank 0 publishes it through a file), and the exchange lists the C++ Halo is built from equal the Python
    mirror's."""
    import re
    import subprocess
    import benchmarks_b200 as b
    from benchmarks_b200.dist import exchange_lists_local
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    drv = os.path.join(root, "benchmarks_b200", "drivers")
    exe = os.path.join(drv, "comm_check")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", drv, "comm_check"], check=True)
    port = _free_port()
    p = 3
    procs = []
    for r in reversed(range(world)):  # rank 0 last: the others must wait for its file
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_PORT=str(port))
        procs.append((r, subprocess.Popen([exe, str(p)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    outs = {}
    for r, pr in procs:
        out, err = pr.communicate(timeout=120)
        assert pr.returncode == 0, err
        outs[r] = out
    ids = {re.search(r"id=([0-9a-f]+)", outs[r]).group(1) for r in range(world)}
    assert len(ids) == 1
    sub, p1, p2 = (2, 1, 1), (-1.0, -1.0, -1.0), (2.8, 0.9, 0.9)
    for r in range(world):
        meshes = {"box": b.BoxMesh(sub, 1, p, p1=p1, p2=p2, n_ranks=world, rank=r),
                  "hang": b.HangingBoxMesh(sub, 1, p, (1, 0, 0), (3, 1, 2), p1=p1, p2=p2, n_ranks=world, rank=r)}
        for name, mesh in meshes.items():
            line = next(l for l in outs[r].splitlines() if l.startswith(name + " "))
            L = exchange_lists_local(mesh)
            want = (f"{name} n_owned={mesh.n_owned} n_ghost={mesh.n_ghost} n_peers={len(L['peers'])} n_send={len(L['send_indices'])} "
                    f"send_sum={int(L['send_indices'].astype(np.int64).sum())}"
                    + "".join(f" peer{int(t)}:recv={int(c)},send={int(s)}" for t, c, s in zip(L["peers"], L["recv_count"], L["send_count"])))
            assert line == want
