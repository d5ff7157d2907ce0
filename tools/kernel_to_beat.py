#!/usr/bin/env python
"""The reference's own CUDA kernels (compiled in place, oracle/_ref/libref_gpu_*.so) against the product's E-vector kernels on
the same B200 and the same device arrays: BK1/BK3/BK5, p = 1..8, ~1e7 DoFs (BASELINE config C2), FP64; plus the DMMA study
(SURVEY K9).  usage: python tools/kernel_to_beat.py --json gpurun_out/r02_kernel_to_beat.json"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchmarks_b200 as b  # noqa: E402
from oracle import ref_gpu  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dofs", type=float, default=1e7)
    ap.add_argument("--ntests", type=int, default=10)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    rows = ref_gpu.kernel_to_beat(b, args.dofs, ntests=args.ntests)
    print(f"{'kind':>5}{'p':>3}{'nelmt':>9}{'ref_us':>10}{'ref GDoF/s':>12}{'ours_us':>10}{'ours GDoF/s':>12}{'x':>7}{'max rel diff':>14}")
    for r in rows:
        d = "-" if r.get("max_rel_diff") is None else f"{r['max_rel_diff']:.1e}"
        print(f"{r['kind']:>5}{r['p']:>3}{r['nelmt']:>9}{1e3*r['ref_ms']:>10.1f}{r['ref_gdofs']:>12.2f}{1e3*r['ours_ms']:>10.1f}{r['ours_gdofs']:>12.2f}{r['speedup']:>7.2f}{d:>14}")
    dm = ref_gpu.dmma_study(b, ntests=args.ntests)
    for r in dm:
        print(f"K9 nq={r['nq']}: DMMA {r['dmma_gdofs']:.2f}  CUDA-core warp {r['cuda_core_warp_gdofs']:.2f}  ours {r['ours_gdofs']:.2f} GDoF/s; "
              f"dmma vs cuda-core {r['dmma_vs_cuda_core_max_rel_diff']:.1e}, vs ours {r['dmma_vs_ours_max_rel_diff']:.1e}")
    if args.json:
        json.dump(dict(lib=b.LIB_PATH, kernel_to_beat=rows, dmma_study=dm), open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
