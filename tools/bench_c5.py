#!/usr/bin/env python
"""BASELINE config C5: CEED BP6 (vector Laplacian, 3 components, GLL collocated) at p = 8 on a smoothly deformed
MappingQ2 mesh with one level of hanging nodes, STRONG scaling over 1/2/4/8 GPUs of one node (the global mesh is fixed,
the p4est curve is cut into world_size pieces).  Not the headline (bench.py is); prints one JSON line on rank 0.

  python tools/bench_c5.py [--cells-log2 5] [--p 8] [--refine-frac 2] [--its 50] [--steps 3]
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_c5.py --cells-log2 6

--cells-log2 n: (2^n)^2 x 2^(n-1) cells before refinement (n = 6: the 64 x 64 x 32 cells of SURVEY section 8d);
--refine-frac f: the corner block of cells/f per axis is refined once (f = 2: an octant, f = 4: 1/64 of the cells).
Timing: CUDA events around `steps` solves of `its` CG iterations, barrier + synchronize on both sides, max over ranks.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchmarks_b200 as b  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells-log2", type=int, default=5)
    ap.add_argument("--p", type=int, default=8)
    ap.add_argument("--refine-frac", type=int, default=2)
    ap.add_argument("--its", type=int, default=50)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--constraints", default="faces", choices=["rows", "faces"])
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    gloo = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
        gloo = dist.new_group(backend="gloo")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    n, p, nc = args.cells_log2, args.p, 3
    cells = (1 << n, 1 << n, 1 << (n - 1))
    hi = tuple(max(c // args.refine_frac, 1) for c in cells)
    t0 = time.perf_counter()
    mesh = b.HangingBoxMesh((2, 2, 1), n - 1, p, (0, 0, 0), hi, n_ranks=world, rank=rank)
    halo = None
    if world > 1:
        from benchmarks_b200.dist import Halo
        halo = Halo(mesh, group=gloo)
    A = b.LaplaceOperator(mesh, quad="gll", p_geo=2, deform=(0.05, 2.0), halo=halo, with_jxw=True, constraints=args.constraints)
    rhs = A.compute_rhs().repeat(nc)
    nloc = mesh.n_owned + mesh.n_ghost
    x = torch.zeros(nc * nloc, dtype=torch.float64, device=dev)
    t_setup = time.perf_counter() - t0
    ctl = b.ReductionControl(args.its, 0.0, 0.0)
    solver = b.SolverCG(ctl, check_every=1 << 30)

    def solve():
        try:
            solver.solve(A, x, rhs, n_components=nc)
        except b.NoConvergence:
            pass  # fixed iteration count per step

    for _ in range(args.warmup):
        solve()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        solve()
    e1.record()
    barrier()
    t_cg = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    # apply only
    src = torch.rand(nc * nloc, dtype=torch.float64, device=dev)
    dst = torch.empty_like(src)
    for _ in range(3):
        A.vmult_components(dst, src, nc)
    barrier()
    e0.record()
    reps = 10
    for _ in range(reps):
        A.vmult_components(dst, src, nc)
    e1.record()
    barrier()
    t_apply = max_over_ranks(e0.elapsed_time(e1) * 1e-3 / reps)
    n_dofs = nc * int(mesh.n_dofs_global)
    if rank == 0:
        print(json.dumps({
            "metric": "BP6 CG GDoF/s (3-component DoFs x iterations / s)", "value": 1e-9 * n_dofs * args.its * args.steps / t_cg,
            "unit": "GDoF/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_cg / args.steps,
            "higher_is_better": True, "scaling": "strong", "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"CEED BP6 vector Laplacian GLL, p={p}, {cells[0]}x{cells[1]}x{cells[2]} cells with the corner block "
                                   f"{hi[0]}x{hi[1]}x{hi[2]} refined once (hanging nodes), deformed MappingQ2 mesh, {args.its} CG iterations per step",
                       "cells": int(mesh.n_cells_global), "n_dofs": n_dofs, "hanging_rows_rank0": int(len(mesh.hang_dof)), "constraints": args.constraints, "setup_s": t_setup},
            "apply_only": {"gdofs": 1e-9 * n_dofs / t_apply, "ms": 1e3 * t_apply}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
