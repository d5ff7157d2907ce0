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


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    gloo = dist.new_group(backend="gloo")
    blocks = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
    ok = True
    for p, quad, nq, nref, overlap in ((2, "gauss", 4, 3, False), (4, "gauss", 6, 3, True), (6, "gll", 7, 2, True), (3, "gauss", 4, 2, True)):
        mesh = b.BoxMesh(blocks, nref, p, n_ranks=world, rank=rank)
        halo = Halo(mesh, group=gloo)
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
        good = err <= 1e-12 and abs(ctl.last_step() - ctl1.last_step()) <= 1 and xerr <= 1e-6
        ok &= bool(good)
        print(f"[rank {rank}/{world}] p={p} {quad} nq={nq} overlap={overlap}: vmult rel err {err:.2e}, CG its {ctl.last_step()} vs {ctl1.last_step()} (1 GPU), x rel err {xerr:.1e} -> {'OK' if good else 'FAIL'}", flush=True)
        del A, A1, halo
    t = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_CHECK", "PASS" if t.item() == 1 else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1 else 1)


if __name__ == "__main__":
    main()
