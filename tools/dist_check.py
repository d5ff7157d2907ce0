#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU): distributed vmult and CG over the
NCCL halo exchange against the same problem solved on one GPU by rank 0 (whose single-GPU path is
itself parity-tested against the CPU oracle in tests/test_operator_gpu.py).
  torchrun --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchmarks_b200 as b  # noqa: E402
from benchmarks_b200.dist import Halo  # noqa: E402


TRANSPORT = "p2p"   # set by main(): every halo of a pass uses this transport


def make_halo(mesh, gloo):
    halo = Halo(mesh, group=gloo)
    if TRANSPORT == "p2p" and not halo.p2p_available():
        raise RuntimeError("P2P windows not available on this box")
    halo.set_transport(TRANSPORT)
    return halo


def lattice_of_local(mesh, A):
    """Global lattice index (Z*ny + Y)*nx + X of every local DoF that an owned cell touches."""
    p, nm = mesh.p, mesh.p + 1
    n = [c * p + 1 for c in mesh.cells]
    idx = mesh.dof_indices.astype(np.int64)
    l = np.arange(nm ** 3)
    a, bb, c = l % nm, (l // nm) % nm, l // (nm * nm)
    xyz = mesh.cell_xyz.astype(np.int64)
    lin = ((xyz[:, 2:3] * p + c[None]) * n[1] + (xyz[:, 1:2] * p + bb[None])) * n[0] + (xyz[:, 0:1] * p + a[None])
    out = np.full(mesh.n_owned + mesh.n_ghost, -1, dtype=np.int64)
    valid = idx != 0xFFFFFFFF
    out[idx[valid]] = lin[valid]
    return out, int(np.prod(n))


def keys_of_local(mesh):
    """Numbering-independent identity of every local DoF of a HangingBoxMesh that an owned cell touches (convention H1 of
    csrc/hangmesh.cc): coarse-level DoFs by their point on the coarse lattice, fine-level DoFs (everything of a child cell
    that is not a vertex of the coarse mesh) by their point on the fine lattice, offset by the coarse lattice size."""
    p, nm = mesh.p, mesh.p + 1
    n0 = [c * p + 1 for c in mesh.cells]
    n1 = [2 * c * p + 1 for c in mesh.cells]
    l = np.arange(nm ** 3)
    a = np.stack([l % nm, (l // nm) % nm, l // (nm * nm)], axis=1)              # [nm^3, 3]
    lvl = mesh.cell_lxyz[:, 0].astype(np.int64)
    F = mesh.cell_lxyz[:, None, 1:].astype(np.int64) * p + a[None]                # lattice point at the cell's own level
    coarse_vertex = (F % (2 * p) == 0).all(axis=2)
    as_coarse = (lvl[:, None] == 0) | coarse_vertex
    C = np.where((lvl == 0)[:, None, None], F, F // 2)                            # coarse lattice point where it applies
    key0 = (C[..., 2] * n0[1] + C[..., 1]) * n0[0] + C[..., 0]
    key1 = int(np.prod(n0)) + (F[..., 2] * n1[1] + F[..., 1]) * n1[0] + F[..., 0]
    key = np.where(as_coarse, key0, key1)
    out = np.full(mesh.n_owned + mesh.n_ghost, -1, dtype=np.int64)
    idx = mesh.dof_indices.astype(np.int64)
    valid = idx != 0xFFFFFFFF
    out[idx[valid]] = key[valid]
    return out, int(np.prod(n0)) + int(np.prod(n1))


def check_hanging(rank, world, gloo):
    """Distributed C^T A C (hanging-node rows whose parents are ghosts, condensation before compress) and the 3-component CG
    against the same two-level mesh on one GPU."""
    blocks = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
    ok = True
    for p, quad, nref, form in ((2, "gauss", 2, "faces"), (4, "gll", 2, "rows"), (7, "gll", 1, "faces")):
        cells = [s << nref for s in blocks]
        lo, hi = (0, 0, 0), tuple(max(c // 2, 1) for c in cells)
        mesh = b.HangingBoxMesh(blocks, nref, p, lo, hi, n_ranks=world, rank=rank)
        halo = make_halo(mesh, gloo)
        kw = dict(quad=quad, deform=(0.03, 1.5), p_geo=2, constraints=form)
        A = b.LaplaceOperator(mesh, halo=halo, **kw)
        m1 = b.HangingBoxMesh(blocks, nref, p, lo, hi)
        A1 = b.LaplaceOperator(m1, **kw)
        key, n_key = keys_of_local(mesh)
        key1, _ = keys_of_local(m1)
        field = np.random.default_rng(9).standard_normal(n_key)
        own = key[: mesh.n_owned]
        sel = own >= 0
        src = A.initialize_dof_vector()
        src[: mesh.n_owned] = torch.from_numpy(np.where(sel, field[np.maximum(own, 0)], 0.0)).cuda()
        dst = A.initialize_dof_vector()
        A.vmult(dst, src)
        s1 = A1.initialize_dof_vector()
        s1[:] = torch.from_numpy(np.where(key1 >= 0, field[np.maximum(key1, 0)], 0.0)).cuda()
        d1 = A1.initialize_dof_vector()
        A1.vmult(d1, s1)
        full = np.zeros(n_key)
        full[key1[key1 >= 0]] = d1.cpu().numpy()[key1 >= 0]
        err = np.abs(dst[: mesh.n_owned].cpu().numpy()[sel] - full[own[sel]]).max() / np.abs(full).max()
        # BP6-style CG: three components, rhs = int phi in each
        nloc, nloc1 = mesh.n_owned + mesh.n_ghost, m1.n_owned
        batch_same = True
        if TRANSPORT == "nccl":  # component-batched ghost update / compress against one exchange per component
            s3 = torch.from_numpy(np.random.default_rng(11 + rank).standard_normal(3 * nloc)).cuda()
            outs = []
            for flag in ("1", "0"):
                os.environ["B200FE_HALO_BATCH"] = flag
                for c in range(3):
                    s3[c * nloc + mesh.n_owned:(c + 1) * nloc] = 0.0
                d3 = torch.zeros_like(s3)
                A.vmult_components(d3, s3.clone(), 3)
                outs.append(d3.clone())
            os.environ["B200FE_HALO_BATCH"] = "1"
            # (atomics reorder the sums inside the cell kernel, so the applies agree to rounding, not bitwise)
            batch_same = bool(((outs[0] - outs[1]).abs().max() <= 1e-12 * outs[1].abs().max()).item())
        rhs = A.compute_rhs().repeat(3)
        x = torch.zeros(3 * nloc, dtype=torch.float64, device="cuda")
        ctl = b.ReductionControl(20000, 1e-16, 1e-9)
        b.SolverCG(ctl).solve(A, x, rhs, n_components=3)
        rhs1 = A1.compute_rhs().repeat(3)
        x1 = torch.zeros(3 * nloc1, dtype=torch.float64, device="cuda")
        ctl1 = b.ReductionControl(20000, 1e-16, 1e-9)
        b.SolverCG(ctl1).solve(A1, x1, rhs1, n_components=3)
        xs = np.zeros(n_key)
        xs[key1[key1 >= 0]] = x1[2 * nloc1:].cpu().numpy()[key1 >= 0]
        xerr = np.abs(x[2 * nloc: 2 * nloc + mesh.n_owned].cpu().numpy()[sel] - xs[own[sel]]).max() / np.abs(xs).max()
        halo.status()
        good = err <= 1e-12 and abs(ctl.last_step() - ctl1.last_step()) <= 1 and xerr <= 1e-6 and batch_same
        ok &= bool(good)
        print(f"[rank {rank}/{world}] [{TRANSPORT}] batched==per-component {batch_same}; hanging p={p} {quad} constraints={form}: {len(mesh.hang_dof)} rows, vmult rel err {err:.2e}, BP6 CG its {ctl.last_step()} vs "
              f"{ctl1.last_step()} (1 GPU), x rel err {xerr:.1e} -> {'OK' if good else 'FAIL'}", flush=True)
        del A, A1, halo
    return ok


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    gloo = dist.new_group(backend="gloo")
    global TRANSPORT
    probe = Halo(b.BoxMesh({1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world], 1, 1, n_ranks=world, rank=rank), group=gloo)
    transports = (["p2p"] if probe.p2p_available() else []) + ["nccl"]
    del probe
    if rank == 0:
        print("transports:", transports, flush=True)
    total = 1
    for TRANSPORT in transports:
        total = min(total, one_pass(rank, world, gloo))
    dist.destroy_process_group()
    sys.exit(0 if total == 1 else 1)


def one_pass(rank, world, gloo):
    blocks = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
    ok = True
    for p, quad, nq, nref, overlap in ((2, "gauss", 4, 3, False), (4, "gauss", 6, 3, True), (6, "gll", 7, 2, True), (3, "gauss", 4, 2, True)):
        mesh = b.BoxMesh(blocks, nref, p, n_ranks=world, rank=rank)
        halo = make_halo(mesh, gloo)
        A = b.LaplaceOperator(mesh, nq=nq, quad=quad, halo=halo, overlap=overlap, deform=(0.03, 1.5), p_geo=2)
        lat, n_lat = lattice_of_local(mesh, A)
        # a global field defined on the lattice so every rank can evaluate its part
        gen = np.random.default_rng(5)
        field = gen.standard_normal(n_lat)
        src = A.initialize_dof_vector()
        own = lat[: mesh.n_owned]
        vals = np.where(own >= 0, field[np.maximum(own, 0)], 0.0)
        src[: mesh.n_owned] = torch.from_numpy(vals).cuda()
        dst = A.initialize_dof_vector()
        A.vmult(dst, src)
        # reference: the whole mesh on this GPU (single rank)
        m1 = b.BoxMesh(blocks, nref, p)
        A1 = b.LaplaceOperator(m1, nq=nq, quad=quad, deform=(0.03, 1.5), p_geo=2)
        lat1, _ = lattice_of_local(m1, A1)
        s1 = A1.initialize_dof_vector()
        s1[:] = torch.from_numpy(np.where(lat1 >= 0, field[np.maximum(lat1, 0)], 0.0)).cuda()
        d1 = A1.initialize_dof_vector()
        A1.vmult(d1, s1)
        full = np.zeros(n_lat)
        full[lat1[lat1 >= 0]] = d1.cpu().numpy()[lat1 >= 0]
        mine = dst[: mesh.n_owned].cpu().numpy()
        sel = own >= 0
        err = np.abs(mine[sel] - full[own[sel]]).max() / np.abs(full).max()
        # CG with the bp3 protocol: iteration counts must agree
        rhs, x = A.compute_rhs(), A.initialize_dof_vector()
        ctl = b.ReductionControl(20000, 1e-16, 1e-9)
        b.SolverCG(ctl).solve(A, x, rhs)
        rhs1, x1 = A1.compute_rhs(), A1.initialize_dof_vector()
        ctl1 = b.ReductionControl(20000, 1e-16, 1e-9)
        b.SolverCG(ctl1).solve(A1, x1, rhs1)
        xs = np.zeros(n_lat)
        xs[lat1[lat1 >= 0]] = x1.cpu().numpy()[lat1 >= 0]
        xerr = np.abs(x[: mesh.n_owned].cpu().numpy()[sel] - xs[own[sel]]).max() / np.abs(xs).max()
        halo.status()
        good = err <= 1e-12 and abs(ctl.last_step() - ctl1.last_step()) <= 1 and xerr <= 1e-6
        ok &= bool(good)
        print(f"[rank {rank}/{world}] [{TRANSPORT}] p={p} {quad} nq={nq} overlap={overlap}: vmult rel err {err:.2e}, CG its {ctl.last_step()} vs {ctl1.last_step()} (1 GPU), x rel err {xerr:.1e} -> {'OK' if good else 'FAIL'}", flush=True)
        del A, A1, halo
    t = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"DIST_CHECK [{TRANSPORT}]", "PASS" if t.item() == 1 else "FAIL", flush=True)
    th = torch.tensor([int(check_hanging(rank, world, gloo))], device="cuda")
    dist.all_reduce(th, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"DIST_CHECK_HANGING [{TRANSPORT}]", "PASS" if th.item() == 1 else "FAIL", flush=True)
    return int(torch.minimum(t, th).item())


if __name__ == "__main__":
    main()
