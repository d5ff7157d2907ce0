#!/usr/bin/env python
"""E-vector kernel sweep: BK1/BK3/BK5, p = 1..8, ~1e7 DoFs (BASELINE config C2), FP64.
Prints the reference drivers' table columns (CEED_BK/include/benchmark_printer.hpp) plus the
fraction of the measured HBM peak.  GDoF/s = nelmt*nm^3/time, bw uses the reference's formulas
(CEED_BK/src/BK{1,3,5}/templated_cuda_benchmark.cc:102,113,102)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchmarks_b200 as b  # noqa: E402


def peak_gbs():
    try:
        return json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dofs", type=float, default=1e7)
    ap.add_argument("--kinds", default="bk1,bk3,bk5")
    ap.add_argument("--degrees", default="1,2,3,4,5,6,7,8")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    peak = peak_gbs()
    rows = []
    print(f"{'kernel':>6}{'p':>3}{'nelmt':>9}{'epb':>5}{'blocks':>7}{'thr':>5}{'DOF':>10}{'time_us':>10}{'GDOF/s':>9}{'GB/s':>9}{'%hbm':>7}")
    for kind in args.kinds.split(","):
        for p in [int(x) for x in args.degrees.split(",")]:
            nm = p + 1
            nq = nm if kind == "bk5" else p + 2
            nelmt = int(args.dofs) // nm ** 3
            ndof, nquad = nelmt * nm ** 3, nelmt * nq ** 3
            basis = np.cos(np.arange(nq * nm, dtype=np.float64))
            dbasis = np.cos(np.arange(nq * nq, dtype=np.float64))
            u = torch.rand(ndof, dtype=torch.float64, device="cuda")
            out = torch.empty_like(u)
            if kind == "bk1":
                J = torch.rand(nquad, dtype=torch.float64, device="cuda")
                f = lambda: b.bk1_apply(p, nq, basis, J, u, out)
                nbytes = 8 * (2 * ndof + nquad)
            else:
                G = torch.rand(6 * nquad, dtype=torch.float64, device="cuda")
                f = (lambda: b.bk5_apply(p, dbasis, G, u, out)) if kind == "bk5" else (lambda: b.bk3_apply(p, nq, basis, dbasis, G, u, out))
                nbytes = 8 * (2 * ndof + 6 * nquad)
            for _ in range(3):
                f()
            torch.cuda.synchronize()
            ts = []
            for _ in range(args.reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); f(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e-3)
            t = float(np.median(ts))
            li = b.bk_launch_info({"bk1": 1, "bk3": 3, "bk5": 5}[kind], p, nq, nelmt)
            row = dict(kernel=kind, p=p, nelmt=nelmt, dof=ndof, time_s=t, t_min=min(ts), gdofs=1e-9 * ndof / t, gbs=1e-9 * nbytes / t,
                       frac=1e-9 * nbytes / t / peak, **li)
            rows.append(row)
            print(f"{kind:>6}{p:>3}{nelmt:>9}{li['elems_per_block']:>5}{li['num_blocks']:>7}{li['threads_per_block']:>5}{ndof:>10}{t*1e6:>10.1f}"
                  f"{row['gdofs']:>9.2f}{row['gbs']:>9.1f}{100*row['frac']:>7.1f}")
            del u, out
            torch.cuda.empty_cache()
    if args.json:
        json.dump(dict(peak_gbs=peak, rows=rows), open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
