#!/usr/bin/env python
"""Operator-apply sweep over degrees for one build of the library (B200FE_LIB selects a tuning variant):
BP5 (GLL collocated), BP3 (QGauss(p+2)) and "bp35" (QGauss(p+1)) at ~1.2e7 DoFs, fraction of the measured HBM roofline.
  B200FE_LIB=$PWD/benchmarks_b200/variants/libb200fe_eo.so python tools/op_sweep.py --json gpurun_out/sweep_eo.json"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchmarks_b200 as b  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--degrees", default="1,2,3,4,5,6,7,8")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--json", default=None)
    ap.add_argument("--mass", action="store_true", help="also BP1 (mass, QGauss(p+2)) and the bp5_kokkos Helmholtz operator")
    ap.add_argument("--geometry", default="stored", help="stored | affine (on-the-fly; the cube cells of these meshes take the separable kernels)")
    args = ap.parse_args()
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    rows = []
    for p in [int(x) for x in args.degrees.split(",")]:
        def ndofs(c):
            n, rem = c // 3, c % 3
            return float(np.prod([((2 if d < rem else 1) << n) * p + 1 for d in range(3)]))
        best = min(range(0, 27), key=lambda c: abs(np.log(ndofs(c) / 1.2e7)))
        mesh = b.BoxMesh.bp3_cycle(best, p)
        for name, kw in (("bp5", dict(quad="gll")), ("bp3", dict(quad="gauss", nq=p + 2)), ("bp35", dict(quad="gauss", nq=p + 1)),
                         ("bp1", dict(quad="gauss", nq=p + 2, kind="mass")), ("helm", dict(quad="gauss", nq=p + 1, kind="helmholtz"))):
            if name in ("bp1", "helm") and not args.mass:
                continue
            op = b.LaplaceOperator(mesh, with_jxw=name in ("bp1", "helm") and args.geometry == "stored", geometry=args.geometry, **kw)
            src = torch.rand(mesh.n_owned, dtype=torch.float64, device="cuda")
            dst = torch.empty_like(src)
            for _ in range(3):
                op.vmult(dst, src)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                op.vmult(dst, src)
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1) * 1e-3 / args.reps
            rows.append(dict(op=name, p=p, n_dofs=int(mesh.n_dofs_global), ms=1e3 * t, gdofs=1e-9 * mesh.n_dofs_global / t,
                             frac=1e-9 * op.algorithmic_bytes() / t / peak, regs=op.launch_info()["regs_per_thread"],
                             blocks_per_sm=op.launch_info()["blocks_per_sm"]))
            print(f"{name:5s} p={p} {rows[-1]['gdofs']:7.2f} GDoF/s  {100 * rows[-1]['frac']:5.1f} % of HBM  regs {rows[-1]['regs']} x {rows[-1]['blocks_per_sm']} CTAs/SM", flush=True)
            del op, src, dst
            torch.cuda.empty_cache()
    if args.json:
        json.dump(dict(lib=b.LIB_PATH, peak_gbs=peak, rows=rows), open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
