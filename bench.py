#!/usr/bin/env python
"""bench.py -- headline benchmark of the b200fe hot path.

Workload (BASELINE.json configs[2], the single-GPU configuration the metric is quoted on):
  CEED BP5 -- collocated GLL Laplacian, p = 6, 64^3 cells per GPU (57,066,625 DoFs on one GPU),
  conjugate gradients from x0 = 0 on rhs = int phi (bp3.cc:184-239), a fixed number of iterations per
  solve (bp5_kokkos/benchmark.cc:355 caps CG at 100 iterations the same way).
  One *step* = one CG solve of `--its` iterations.  metric = GDoF/s = global DoFs * iterations / time
  (bp5_kokkos/benchmark.cc:406).  Weak scaling: 64^3 cells per GPU, box of 2x1x1 / 2x2x1 / 2x2x2 blocks.

Contract keys: value (device-resident vectors, CUDA events, max over ranks), e2e (same solve through
the host-buffer C-ABI entry point b200fe_cg_solve_host: H2D of b and D2H of x inside the timed
region), roofline (the sum-factorised cell kernel, timed per launch with CUDA events inside the
library during the timed region), cpu_baseline (the oracle's C+OpenMP port on the host cores, on a
bounded sample), clocks, gpu_launches.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  torchrun --nproc-per-node N ... bench.py --gpus N ...
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P_DEGREE = 6
CELLS_LOG2 = 6          # 64^3 cells per GPU
CPU_SAMPLE_LOG2 = 5     # CPU arm: 32^3 cells (7,189,057 DoFs), same operator and solver
BLOCKS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa(local_rank):
    """Pin this rank to the host cores NVML reports as local to its GPU (before any pinned allocation), so that the e2e
    arm's pinned buffers are first-touched on the GPU's own NUMA node: with 8 ranks the H2D / D2H copies of a step
    (3.65 GB each way) otherwise all go through whichever node the launcher started on.  Returns the number of cores or None."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = "GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        n = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (n + 63) // 64)
        cpus = {i for i in range(n) if (mask[i // 64] >> (i % 64)) & 1} & os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def colors_of(cell_xyz):
    col = (cell_xyz[:, 0] & 1) + 2 * (cell_xyz[:, 1] & 1) + 4 * (cell_xyz[:, 2] & 1)
    order = np.argsort(col, kind="stable")
    off = np.concatenate([[0], np.cumsum(np.bincount(col, minlength=8))])
    return off.astype(np.uint32), order.astype(np.uint32)


def host_cores():
    """Cores this process may run on (torchrun exports OMP_NUM_THREADS=1 to its ranks, which must not shrink the CPU arm)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_cg_arm(p, cells_log2, its, repeats=1):
    """The oracle's C + OpenMP restatement of the same operator + CG on all host cores (kind 'port': the reference's own
    CPU path, bk3_dealii / deal.II, cannot be built here -- DESIGN.md).  Stands on oracle/ alone: mesh, numbering,
    Dirichlet mask, right-hand side and solver all come from the checker, nothing from the product package."""
    import oracle
    fe = oracle.fe
    oracle.port.set_num_threads(host_cores())
    mesh = fe.BoxMesh((1, 1, 1), cells_log2)
    rd = fe.rank_data_single_fast(mesh, p)
    bas = fe.basis_1d(p, p + 1, "gll")
    h = mesh.h[0]
    w = bas["wq"]
    nq = p + 1
    W3 = np.einsum("r,q,p->rqp", w, w, w).ravel()
    G = np.zeros((mesh.n_cells, 6, nq ** 3))
    G[:, 0] = G[:, 3] = G[:, 5] = h * W3                 # cube cells: G = diag(h w_q)
    # b = int phi (bp3.cc:208-224), the GPU arm's right-hand side; collocated basis: the cell vector is JxW itself
    # (same as fe.rhs_one(rd, bas, JxW) with B = I, without its dense einsum)
    idx = rd["dof_indices"]
    valid = idx != fe.INVALID
    loc = np.broadcast_to(h ** 3 * W3, idx.shape)
    rhs = np.bincount(idx[valid].astype(np.int64), weights=loc[valid], minlength=rd["n_owned"]).astype(np.float64)
    kw = dict(nm=p + 1, nq=nq, collocated=True, flags=1, shape_values=bas["B"].T.copy(), co_shape_gradients=bas["D"].T.copy(),
              G=G, JxW=None, dof_indices=rd["dof_indices"], colors=colors_of(mesh.cell_xyz), constrained=rd["constrained"],
              max_it=its, abs_tol=0.0, rel_tol=0.0)
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        _, n_it, _, _, _ = oracle.port.cg_solve(rhs, **kw)
        times.append(time.perf_counter() - t0)
        assert n_it == its
    return dict(n_dofs=int(rd["n_owned"]), its=its, times=times, cores=oracle.port.num_threads())


def serial_kernel_arm(budget_s=2.0):
    """the reference's OWN serial BK1/BK3/BK5 kernels (oracle/_ref, compiled verbatim from
    /root/reference, single-threaded by construction) timed on one host core, p = 1..8, nelmt sized for ~budget_s/24 each."""
    import oracle
    if oracle.ref is None:
        return None
    rows = []
    for kind in ("bk1", "bk3", "bk5"):
        for p in range(1, 9):
            nm = p + 1
            nq = nm if kind == "bk5" else p + 2
            nelmt = max(8, int(2.0e5 / nq ** 3))
            k = oracle.kat_inputs(kind, p, nelmt)
            rng = np.random.default_rng(p)
            u = rng.uniform(-1, 1, k["u"].size)
            t0 = time.perf_counter()
            reps = 0
            while True:
                if kind == "bk1":
                    oracle.ref.bk1(nq, k["basis"], k["JxW"], u)
                elif kind == "bk3":
                    oracle.ref.bk3(nq, k["basis"], k["dbasis"], k["G"], u)
                else:
                    oracle.ref.bk5(nq, k["dbasis"], k["G"], u, which="ceedbk")
                reps += 1
                dt = time.perf_counter() - t0
                if dt > budget_s / 24 or reps >= 50:
                    break
            rows.append({"kind": kind, "p": p, "nelmt": nelmt, "gdofs_per_core": 1e-9 * nelmt * nm ** 3 * reps / dt})
    return rows


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    its = args.cpu_its
    r = cpu_cg_arm(P_DEGREE, CPU_SAMPLE_LOG2, its, repeats=args.warmup + args.steps)
    t = float(np.mean(r["times"][args.warmup:]))
    val = 1e-9 * r["n_dofs"] * its / t
    sample = (f"BP5 p={P_DEGREE} CG, {2**CPU_SAMPLE_LOG2}^3 cells ({r['n_dofs']} DoFs), {its} iterations per step, rhs = int phi, x0 = 0; "
              f"oracle/fe_oracle.c (C + OpenMP, 8 cells side by side) on {r['cores']} threads")
    serial = serial_kernel_arm()
    out = {
        "impl": "reference", "metric": "BP5 CG GDoF/s (DoFs x iterations / s)", "value": val, "unit": "GDoF/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"CEED BP5 collocated GLL Laplacian CG, p={P_DEGREE}, 64^3 cells/GPU; CPU arm on a bounded sample", "sample": sample},
        "cpu_baseline": {"value": val, "unit": "GDoF/s", "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "GDoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}
    if serial:
        out["cpu_baseline_serial_kernels"] = {"kind": "reference", "cores": 1, "unit": "GDoF/s per core",
                                              "what": "CEED_BK serial BK1/BK3/BK5 kernels compiled verbatim (oracle/_ref), E-vector DoFs x applications / s",
                                              "rows": serial}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--its", type=int, default=100, help="CG iterations per step")
    ap.add_argument("--cpu-its", type=int, default=100, help="CG iterations per step of the CPU arm (same count as the GPU arm)")
    ap.add_argument("--p", type=int, default=P_DEGREE)
    ap.add_argument("--cells-log2", type=int, default=CELLS_LOG2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--overlap", type=int, default=1)
    ap.add_argument("--transport", default="auto", choices=["auto", "p2p", "nccl"],
                    help="ghost exchange / reductions at N > 1: peer stores over CUDA-IPC windows, or NCCL; auto = p2p when every peer could be mapped")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import benchmarks_b200 as b
    from benchmarks_b200._lib import lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cores = bind_to_gpu_numa(local_rank) if world > 1 else None
    setup_group = None
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner there when NCCL_DEBUG=VERSION
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
        setup_group = dist.new_group(backend="gloo")

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

    p, its = args.p, args.its
    if world not in BLOCKS:
        raise SystemExit(f"--gpus must be one of {sorted(BLOCKS)}")
    t_setup = time.perf_counter()
    mesh = b.BoxMesh(BLOCKS[world], args.cells_log2, p, n_ranks=world, rank=rank)
    halo = None
    if world > 1:
        from benchmarks_b200.dist import Halo
        halo = Halo(mesh, group=setup_group)
        if args.transport != "auto":
            halo.set_transport(args.transport)
    A = b.LaplaceOperator(mesh, quad="gll", halo=halo, overlap=bool(args.overlap))
    exclusive_interior = A.launch_info()["exclusive_interior"]
    rhs = A.compute_rhs()
    x = A.initialize_dof_vector()
    t_setup = time.perf_counter() - t_setup
    n_dofs = int(mesh.n_dofs_global)
    ctl = b.ReductionControl(its, 0.0, 0.0)      # fixed iteration count: tolerances unreachable
    solver = b.SolverCG(ctl, check_every=1 << 30)

    def solve_device():
        try:
            solver.solve(A, x, rhs)
        except b.NoConvergence:
            pass  # expected: the step is `its` iterations (bp5_kokkos/benchmark.cc:370-374)

    h_b = rhs[: mesh.n_owned].cpu().pin_memory()
    h_x = torch.empty(mesh.n_owned, dtype=torch.float64).pin_memory()

    def solve_host():
        try:
            solver.solve_host(A, h_x, h_b)
        except b.NoConvergence:
            pass

    # ---- device-resident arm ------------------------------------------------------------------
    for _ in range(args.warmup):
        solve_device()
    assert ctl.last_step() == its
    A.timing_enable(args.steps * its * 3 + 16)
    launches0 = lib.b200fe_launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        solve_device()
    e1.record()
    barrier()
    t_dev = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    clocks = sampler.stop() if rank == 0 else None
    gpu_launches = int(lib.b200fe_launch_count() - launches0)
    k_ms, k_n = A.timing_read()
    A.timing_enable(0)
    res_final = ctl.last_value() / ctl.initial_value()

    # ---- end-to-end arm: host buffers through b200fe_cg_solve_host ----------------------------
    for _ in range(2):
        solve_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        solve_host()
    barrier()
    t_e2e = max_over_ranks(time.perf_counter() - t0)

    # ---- N > 1: transport (P2P peer stores vs NCCL) x overlap (split cell launch vs one launch), one step each ------
    comm_ab = None
    if world > 1 and not args.no_sweep:
        comm_ab = {}
        def timed_step(op):
            xx, rr = op.initialize_dof_vector(), op.compute_rhs()
            sv = b.SolverCG(b.ReductionControl(its, 0.0, 0.0), check_every=1 << 30)
            def run():
                try:
                    sv.solve(op, xx, rr)
                except b.NoConvergence:
                    pass
            run()
            barrier()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            run()
            c1.record()
            barrier()
            return max_over_ranks(c0.elapsed_time(c1)) / its   # ms per CG iteration
        start = halo.transport()
        A_flat = b.LaplaceOperator(mesh, quad="gll", halo=halo, overlap=False)
        for tr in (["p2p"] if halo.p2p_available() else []) + ["nccl"]:
            halo.set_transport(tr)
            comm_ab[f"{tr}_overlap_ms_per_it"] = timed_step(A)
            comm_ab[f"{tr}_no_overlap_ms_per_it"] = timed_step(A_flat)
        halo.set_transport(start)
        halo.status()
        del A_flat
        torch.cuda.empty_cache()

    # ---- roofline of the dominant kernel ------------------------------------------------------
    peaks, peak_kind = measured_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes = A.algorithmic_bytes()          # per launch over all local cells (single phase)
    # cell-kernel time per operator application on this rank (1 launch, or up to 3 with the overlap split)
    n_applies = args.steps * its
    launches_per_apply = k_n / max(n_applies, 1)
    k_avg_s = 1e-3 * k_ms / max(n_applies, 1)
    achieved = 1e-9 * alg_bytes / k_avg_s if k_n else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"bp5_p{p}_bytes_per_launch")
    except Exception:
        pass

    # ---- apply-only number and degree sweep (explains the headline; not the headline) ---------
    def time_apply(op, reps=20):
        src = op.initialize_dof_vector()
        src[: op.mesh.n_owned] = torch.rand(op.mesh.n_owned, dtype=torch.float64, device=dev)
        dst = op.initialize_dof_vector()
        for _ in range(3):
            op.vmult(dst, src)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(reps):
            op.vmult(dst, src)
        a1.record()
        barrier()
        return max_over_ranks(a0.elapsed_time(a1) * 1e-3 / reps)

    t_apply = time_apply(A)
    apply_info = {"gdofs": 1e-9 * n_dofs / t_apply, "ms": 1e3 * t_apply,
                  "frac_of_hbm_roofline": 1e-9 * alg_bytes / t_apply / peak}
    # ---- "next" row, reported separately: geometric factors evaluated on the fly for the affine
    #      cells of this mesh (six constants per cell instead of 48 nq^3 bytes) -- different algorithmic bytes
    otf = None
    if not args.no_sweep:
        os.environ["B200FE_CARTESIAN"] = "0"   # (read at operator creation) the general affine kernel: parallelepiped cells
        A_otf = b.LaplaceOperator(mesh, quad="gll", halo=halo, overlap=bool(args.overlap), geometry="affine", with_jxw=False)
        del os.environ["B200FE_CARTESIAN"]
        t_otf = time_apply(A_otf)
        otf = {"what": "BP5 operator apply with on-the-fly affine geometry, general (parallelepiped) kernel (not the headline; own byte count)",
               "gdofs": 1e-9 * n_dofs / t_otf, "ms": 1e3 * t_otf, "algorithmic_bytes": A_otf.algorithmic_bytes(),
               "achieved_gbs": 1e-9 * A_otf.algorithmic_bytes() / t_otf, "speedup_vs_stored_G": t_apply / t_otf,
               "frac_of_hbm_roofline_own_bytes": 1e-9 * A_otf.algorithmic_bytes() / t_otf / peak}
        del A_otf
        # the cells of this mesh (and of every mesh of the reference drivers) are axis-aligned boxes: separable kernel
        try:
            A_cart = b.LaplaceOperator(mesh, quad="gll", halo=halo, overlap=bool(args.overlap), geometry="affine", with_jxw=False)
            assert A_cart.launch_info()["cartesian"] == 1
            t_cart = time_apply(A_cart)
            x_c = A_cart.initialize_dof_vector()
            ctl_c = b.ReductionControl(its, 0.0, 0.0)
            solver_c = b.SolverCG(ctl_c, check_every=1 << 30)

            def solve_cart():
                try:
                    solver_c.solve(A_cart, x_c, rhs)
                except b.NoConvergence:
                    pass
            for _ in range(3):
                solve_cart()
            barrier()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(5):
                solve_cart()
            c1.record()
            barrier()
            t_cg_cart = max_over_ranks(c0.elapsed_time(c1) * 1e-3) / 5
            otf["cartesian_cells"] = {
                "what": "the same apply with the separable kernel for axis-aligned cells (deal.II's cartesian cell type; detected on the per-cell constants): "
                        "c_rr SxWxW + c_ss WxSxW + c_tt WxWxS with S = D^T W D, three 1-D contractions per point; cg = the headline CG step with this operator",
                "gdofs": 1e-9 * n_dofs / t_cart, "ms": 1e3 * t_cart, "algorithmic_bytes": A_cart.algorithmic_bytes(),
                "achieved_gbs": 1e-9 * A_cart.algorithmic_bytes() / t_cart, "speedup_vs_stored_G": t_apply / t_cart,
                "frac_of_hbm_roofline_own_bytes": 1e-9 * A_cart.algorithmic_bytes() / t_cart / peak,
                "cg_gdofs": 1e-9 * n_dofs * its / t_cg_cart, "cg_ms_per_step": 1e3 * t_cg_cart,
                "cg_relative_residual_after_step": ctl_c.last_value() / ctl_c.initial_value(),
                "cg_relative_residual_after_step_stored_G": res_final}
            del A_cart, x_c
        except Exception as exc:
            otf["cartesian_cells"] = {"error": repr(exc)}
        # general hexahedra (trilinear cells, the Jacobian rebuilt at every quadrature point): same mesh, vertices displaced
        try:
            A_tri = b.LaplaceOperator(mesh, quad="gll", halo=halo, overlap=bool(args.overlap), geometry="trilinear", with_jxw=False, p_geo=1, deform=(0.02, 1.5))
            t_tri = time_apply(A_tri)
            otf["trilinear_cells"] = {"what": "the same apply on a vertex-displaced (MappingQ1) mesh, G = JxW J^-1 J^-T rebuilt at every quadrature point from 8 vertices per cell",
                                      "gdofs": 1e-9 * n_dofs / t_tri, "ms": 1e3 * t_tri, "algorithmic_bytes": A_tri.algorithmic_bytes(),
                                      "achieved_gbs": 1e-9 * A_tri.algorithmic_bytes() / t_tri, "speedup_vs_stored_G": t_apply / t_tri,
                                      "geometry_bytes_per_cell": 192, "stored_G_bytes_per_cell": 48 * (p + 1) ** 3}
            del A_tri
        except Exception as exc:
            otf["trilinear_cells"] = {"error": repr(exc)}
    sweep = None
    if world == 1 and not args.no_sweep:
        sweep = []
        A = rhs = x = None
        torch.cuda.empty_cache()
        for pp in range(1, 9):
            # ~1e7 DoFs: cells per axis ~ (1e7^(1/3))/p, rounded to the reference's mesh family
            def ndofs(c):  # bp3.cc:443-468
                n, rem = c // 3, c % 3
                return float(np.prod([((2 if d < rem else 1) << n) * pp + 1 for d in range(3)]))
            best = min(range(0, 27), key=lambda c: abs(np.log(ndofs(c) / 1.2e7)))
            m2 = b.BoxMesh.bp3_cycle(best, pp)
            # bp5 / bp3: the metric's operators; bp1 (mass, QGauss(p+2)): BASELINE config C2 "BK1/BP1 mass-matrix apply + CG"
            for name, kw in (("bp5", dict(quad="gll")), ("bp3", dict(quad="gauss", nq=pp + 2)), ("bp1", dict(quad="gauss", nq=pp + 2, kind="mass"))):
                op = b.LaplaceOperator(m2, with_jxw=(name == "bp1"), **kw)
                t = time_apply(op, reps=10)
                row = {"op": name, "p": pp, "n_dofs": int(m2.n_dofs_global), "gdofs": 1e-9 * m2.n_dofs_global / t,
                       "frac_of_hbm_roofline": 1e-9 * op.algorithmic_bytes() / t / peak, "even_odd_kernel": op.launch_info()["even_odd"]}
                if name == "bp1":   # CG on the mass matrix, 20 iterations from x0 = 0, rhs = int phi
                    rb, xb = op.compute_rhs(), op.initialize_dof_vector()
                    cb = b.ReductionControl(20, 0.0, 0.0)
                    sb = b.SolverCG(cb, check_every=1 << 30)
                    for _ in range(2):
                        try:
                            sb.solve(op, xb, rb)
                        except b.NoConvergence:
                            pass
                    torch.cuda.synchronize()
                    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    g0.record()
                    try:
                        sb.solve(op, xb, rb)
                    except b.NoConvergence:
                        pass
                    g1.record()
                    torch.cuda.synchronize()
                    row["cg_gdofs"] = 1e-9 * m2.n_dofs_global * 20 / (g0.elapsed_time(g1) * 1e-3)
                    del rb, xb
                sweep.append(row)
                del op
                torch.cuda.empty_cache()

    # ---- E-vector kernels vs degree at BASELINE config C2 size (1e7 DoFs): BK1 / BK3 / BK5 with the reference drivers' own
    #      cos() test matrices (CEED_BK/src/BK3/templated_cuda_benchmark.cc:50-66; no symmetry -> plain contractions) and, for the
    #      interpolated kernels, with the real Gauss / GLL matrices of FE_Q(p) (SURVEY 8d "strong mode"; even-odd kernel) -----------
    bk_sweeps = None
    if world == 1 and not args.no_sweep:
        bk_sweeps = {"bk1": [], "bk3": [], "bk5": [], "bk1_real_basis": [], "bk3_real_basis": []}
        for kind in ("bk1", "bk3", "bk5"):
            for pp in range(1, 9):
                nm = pp + 1
                nq = nm if kind == "bk5" else pp + 2
                nelmt = 10_000_000 // nm ** 3
                u = torch.rand(nelmt * nm ** 3, dtype=torch.float64, device=dev)
                geo = torch.rand(nelmt * (1 if kind == "bk1" else 6) * nq ** 3, dtype=torch.float64, device=dev)
                out_k = torch.empty_like(u)
                # reference formulas: CEED_BK/src/BK1/templated_cuda_benchmark.cc:102, BK3/...:113, BK5/...:102
                nbytes = 8 * (2 * nelmt * nm ** 3 + (1 if kind == "bk1" else 6) * nelmt * nq ** 3)
                modes = [("", np.cos(np.arange(nq * nm, dtype=np.float64)), np.cos(np.arange(nq * nq, dtype=np.float64)))]
                if kind != "bk5":
                    bas = b.basis_1d(pp, nq, b.QUAD_GAUSS)
                    modes.append(("_real_basis", np.ascontiguousarray(bas["shape_values"].reshape(nm, nq).T),
                                  np.ascontiguousarray(bas["co_shape_gradients"].reshape(nq, nq).T)))
                for suffix, basis, dbasis in modes:
                    if kind == "bk1":
                        f = lambda: b.bk1_apply(pp, nq, basis, geo, u, out_k)
                    elif kind == "bk3":
                        f = lambda: b.bk3_apply(pp, nq, basis, dbasis, geo, u, out_k)
                    else:
                        f = lambda: b.bk5_apply(pp, dbasis, geo, u, out_k)
                    for _ in range(3):
                        f()
                    torch.cuda.synchronize()
                    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    k0.record()
                    for _ in range(10):
                        f()
                    k1.record()
                    torch.cuda.synchronize()
                    t = k0.elapsed_time(k1) * 1e-4
                    bk_sweeps[kind + suffix].append({"p": pp, "nelmt": nelmt, "gdofs": 1e-9 * nelmt * nm ** 3 / t, "gbs": 1e-9 * nbytes / t,
                                                     "frac_of_hbm_roofline": 1e-9 * nbytes / t / peak})
                del u, geo, out_k
                torch.cuda.empty_cache()

    # ---- the kernel to beat: the reference's OWN CUDA kernels (compiled in place for sm_100a, T = double,
    #      the reference drivers' launch shapes; oracle/_ref/libref_gpu_*.so -- bench infrastructure, not the product) timed on the
    #      same device arrays beside ours, and SURVEY K9: its FP64 tensor-core (DMMA) BK1 kernel ------------------------------------
    ktb = None
    if world == 1 and not args.no_sweep:
        try:
            from oracle import ref_gpu
            if ref_gpu.available():
                ktb = {"kernels": [{k: r[k] for k in ("kind", "p", "nelmt", "ref_gdofs", "ours_gdofs", "speedup", "max_rel_diff")}
                                   for r in ref_gpu.kernel_to_beat(b, 1e7, ntests=5)],
                       "dmma_study": ref_gpu.dmma_study(b, ntests=5),
                       "what": "CEED_BK templated_cuda_kernels.cuh BK1/BK3/BK5 <double> at the drivers' default launch shape vs b200fe_bk*_apply, "
                               "1e7 DoFs, best of 5 launches each (CUDA events); max_rel_diff = our output vs the reference kernel's output"}
            else:
                ktb = {"unavailable": "oracle/_ref/libref_gpu_*.so not built (needs /root/reference at build time)"}
        except Exception as exc:
            ktb = {"error": repr(exc)}

    # ---- BASELINE config C5 on one GPU (explains, not the headline): CEED BP6 = vector Laplacian, 3 components, GLL
    #      collocated, p = 8, smoothly deformed MappingQ2 mesh, lower octant refined once (hanging nodes) -------------
    c5 = None
    if world == 1 and not args.no_sweep:
        try:
            hm = b.HangingBoxMesh((1, 1, 1), 5, 8, (0, 0, 0), (16, 16, 16))   # 32^3 cells, 16^3 of them refined: 61,440 cells
            op6 = b.LaplaceOperator(hm, quad="gll", p_geo=2, deform=(0.05, 2.0), with_jxw=False)
            n1 = hm.n_owned
            src6 = torch.rand(3 * n1, dtype=torch.float64, device=dev)
            dst6 = torch.empty_like(src6)
            for _ in range(3):
                op6.vmult_components(dst6, src6, 3)
            torch.cuda.synchronize()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(10):
                op6.vmult_components(dst6, src6, 3)
            c1.record()
            torch.cuda.synchronize()
            t6 = c0.elapsed_time(c1) * 1e-4
            # constraints in the face-structured form (default): the index blocks are read by distribute and by condense
            con_bytes = 4 * (hm.face_parents.size + hm.face_children.size)
            if op6.multi_component_kernel():   # G and the index table of a cell are read once for the three components
                bytes6 = op6.algorithmic_bytes() + 2 * 32 * hm.n_owned + 2 * con_bytes
            else:                              # one cell-kernel launch per component: G streamed three times
                bytes6 = 3 * op6.algorithmic_bytes() + 2 * con_bytes
            c5 = {"what": "BP6 vector Laplacian apply (3 components, GLL, p=8, deformed MappingQ2 mesh, hanging nodes), 1 GPU",
                  "cells": int(hm.n_cells_global), "n_dofs_3_components": 3 * int(hm.n_dofs_global), "hanging_dofs": int(len(hm.hang_dof)), "constraint_face_blocks": int(len(hm.face_parents)),
                  "ms": 1e3 * t6, "gdofs": 1e-9 * 3 * hm.n_dofs_global / t6, "algorithmic_bytes": int(bytes6),
                  "multi_component_kernel": bool(op6.multi_component_kernel()),
                  "frac_of_hbm_roofline": 1e-9 * bytes6 / t6 / peak}
            del op6, src6, dst6, hm
            torch.cuda.empty_cache()
        except Exception as exc:  # never let an explanatory extra take the headline line down
            c5 = {"error": repr(exc)}

    # ---- BASELINE config C4 + parity against the reference's committed goldens, at every N -----------------------------
    # CEED BP3 (QGauss(p+2)) Laplacian CG, p = 4, 64^3 cells per GPU on the reference's box family: the meshes of
    # CEED_bp/results/1xGH200_P4.txt:646-648 and bakeoff_problems_dealii/tests/scaling_bp35_kokkos_kernel/results/4_GPU.out:779
    # (16,974,593 / 33,883,137 / 67,634,433 / 135,005,697 DoFs at N = 1 / 2 / 4 / 8) -> 722 / 1253 / 1389 / 1454 iterations
    # with ReductionControl(1e9, 1e-16, 1e-9), rhs = int phi, x0 = 0 (bp3.cc:266-285).  Iteration counts are
    # machine-independent: this is the multi-GPU correctness signal of the scaling run.
    parity = None
    if not args.no_sweep and args.cells_log2 == CELLS_LOG2:
        try:
            golden = {1: (722, 0.9717, 16974593), 2: (1253, 0.9836, 33883137), 4: (1389, 0.9852, 67634433), 8: (1454, 0.9858, 135005697)}[world]
            A = rhs = x = None   # (the closures above hold cells, not tensors: this frees the BP5 operator's G)
            torch.cuda.empty_cache()
            m4 = mesh if p == 4 else b.BoxMesh(BLOCKS[world], CELLS_LOG2, 4, n_ranks=world, rank=rank)
            h4 = halo
            if world > 1 and p != 4:
                from benchmarks_b200.dist import Halo
                h4 = Halo(m4, group=setup_group)
                if args.transport != "auto":
                    h4.set_transport(args.transport)
            A4 = b.LaplaceOperator(m4, nq=6, quad="gauss", halo=h4, overlap=bool(args.overlap), with_jxw=True)
            rhs4 = A4.compute_rhs()
            x4 = A4.initialize_dof_vector()
            ctl4 = b.ReductionControl(10 ** 9, 1e-16, 1e-9)
            s4 = b.SolverCG(ctl4)
            s4.solve(A4, x4, rhs4)           # warm-up solve (also the parity solve)
            barrier()
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            q0.record()
            s4.solve(A4, x4, rhs4)
            q1.record()
            barrier()
            t_cg4 = max_over_ranks(q0.elapsed_time(q1) * 1e-3)
            its4 = ctl4.last_step()
            red4 = (ctl4.last_value() / ctl4.initial_value()) ** (1.0 / max(its4, 1))
            # true residual of the returned solution, recomputed with one more apply
            r4 = A4.initialize_dof_vector()
            A4.vmult(r4, x4)
            no = m4.n_owned
            num = torch.stack([((rhs4[:no] - r4[:no]) ** 2).sum(), (rhs4[:no] ** 2).sum()])
            if world > 1:
                dist.all_reduce(num)
            true_rel = float(torch.sqrt(num[0] / num[1]))
            t_ap4 = time_apply(A4)
            nd4 = int(m4.n_dofs_global)
            parity = {"what": "CEED BP3 CG, p=4, nq=6, 64^3 cells/GPU, to 1e-9 (BASELINE config C4); golden = reference's committed logs",
                      "n_dofs": nd4, "golden_n_dofs": golden[2], "its": int(its4), "golden": golden[0], "its_ok": abs(its4 - golden[0]) <= 1 and nd4 == golden[2],
                      "cg_reduction": red4, "golden_cg_reduction": golden[1], "true_relative_residual": true_rel,
                      "cg_ms_per_iteration": 1e3 * t_cg4 / max(its4, 1), "cg_gdofs": 1e-9 * nd4 * its4 / t_cg4,
                      "apply_ms": 1e3 * t_ap4, "apply_gdofs": 1e-9 * nd4 / t_ap4,
                      "apply_frac_of_hbm_roofline": 1e-9 * A4.algorithmic_bytes() / t_ap4 / peak,
                      "kernel_even_odd": A4.launch_info()["even_odd"]}
            del A4, x4, r4
            torch.cuda.empty_cache()
            # the same solve with the geometry evaluated on the fly (reported separately): the cube cells of the
            # reference's meshes take the separable kernel on the nodal values -- same golden iteration count expected
            try:
                A4c = b.LaplaceOperator(m4, nq=6, quad="gauss", halo=h4, overlap=bool(args.overlap), with_jxw=False, geometry="affine")
                x4c = A4c.initialize_dof_vector()
                ctl4c = b.ReductionControl(10 ** 9, 1e-16, 1e-9)
                s4c = b.SolverCG(ctl4c)
                s4c.solve(A4c, x4c, rhs4)
                barrier()
                q0.record()
                s4c.solve(A4c, x4c, rhs4)
                q1.record()
                barrier()
                t_cg4c = max_over_ranks(q0.elapsed_time(q1) * 1e-3)
                its4c = ctl4c.last_step()
                t_ap4c = time_apply(A4c)
                parity["on_the_fly_cartesian"] = {
                    "cartesian_kernel": A4c.launch_info()["cartesian"], "its": int(its4c), "its_ok": abs(its4c - golden[0]) <= 1,
                    "cg_ms_per_iteration": 1e3 * t_cg4c / max(its4c, 1), "cg_gdofs": 1e-9 * nd4 * its4c / t_cg4c,
                    "apply_ms": 1e3 * t_ap4c, "apply_gdofs": 1e-9 * nd4 / t_ap4c, "apply_speedup_vs_stored_G": t_ap4 / t_ap4c,
                    "apply_frac_of_hbm_roofline_own_bytes": 1e-9 * A4c.algorithmic_bytes() / t_ap4c / peak}
                del A4c, x4c
            except Exception as exc:
                parity["on_the_fly_cartesian"] = {"error": repr(exc)}
            del rhs4
            torch.cuda.empty_cache()
        except Exception as exc:  # never let an explanatory extra take the headline line down
            parity = {"error": repr(exc)}

    # ---- CPU baseline (rank 0, N = 1 only) ----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_cg_arm(p, CPU_SAMPLE_LOG2, args.cpu_its, repeats=1)
        cpu = {"value": 1e-9 * r["n_dofs"] * r["its"] / r["times"][0], "unit": "GDoF/s", "cores": r["cores"], "kind": "port",
               "sample": f"BP5 p={p} CG, {2**CPU_SAMPLE_LOG2}^3 cells ({r['n_dofs']} DoFs), {r['its']} iterations, oracle/fe_oracle.c (C + OpenMP, 8 cells side by side)"}

    if rank == 0:
        value = 1e-9 * n_dofs * its * args.steps / t_dev
        out = {
            "metric": "BP5 CG GDoF/s (DoFs x iterations / s)", "value": value, "unit": "GDoF/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"CEED BP5 collocated GLL Laplacian CG, p={p}, {2**args.cells_log2}^3 cells per GPU "
                                   f"({n_dofs} DoFs global), {its} CG iterations per step, rhs = int phi, x0 = 0",
                       "mesh_blocks": list(BLOCKS[world]), "n_dofs": n_dofs, "cg_iterations_per_step": its,
                       "l2_policy": "working set (G 4.3 GB + vectors) >> 126 MB L2; no flush needed",
                       "overlap_halo_with_interior_cells": bool(args.overlap) and world > 1,
                       "transport": (halo.transport() if halo is not None else None),
                       "host_cores_bound_to_gpu_numa_node": numa_cores,
                       "relative_residual_after_step": res_final, "setup_s": t_setup},
            "e2e": {"value": 1e-9 * n_dofs * its * args.steps / t_e2e, "unit": "GDoF/s",
                    "h2d_bytes_per_step": int(mesh.n_owned) * 8 * world, "d2h_bytes_per_step": int(mesh.n_owned) * 8 * world,
                    "api": "b200fe_cg_solve_host (pinned host b -> device, CG, device x -> pinned host)"},
            "gpu_launches": gpu_launches,
            "roofline": {"bound": "hbm", "kernel": f"sumfact2_kernel<{p+1},{p+1},collocated,laplace,lvec> (BP5 cell kernel: gather + D^T G D + atomic scatter + fused p.Ap)",
                         "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak if peak else None,
                         "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": 1e3 * k_avg_s,
                         "launches_timed": k_n, "launches_per_apply": launches_per_apply,
                         "exclusive_interior_stores": exclusive_interior, "kernel_share_of_step": (1e-3 * k_ms) / t_dev if t_dev else None},
            "apply_only": apply_info,
            "clocks": clocks,
        }
        if cpu:
            out["cpu_baseline"] = cpu
        if otf:
            out["apply_on_the_fly_affine_geometry"] = otf
        if sweep:
            out["degree_sweep_apply"] = sweep
        if bk_sweeps:
            for k_, v_ in bk_sweeps.items():
                out[f"degree_sweep_{k_}_evector" if not k_.endswith("_real_basis") else f"degree_sweep_{k_[:3]}_evector_real_basis"] = v_
        if comm_ab:
            out["comm_ab"] = comm_ab
        if ktb:
            out["kernel_to_beat"] = ktb
        if parity:
            out["parity"] = parity
        if c5:
            out["bp6_hanging_nodes_p8"] = c5
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
