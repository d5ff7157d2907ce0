#!/bin/bash
# 1 GPU: plane-per-thread mass kernel (BK1 / BP1, p = 2..6): parity suite + A/B against the generic kernel
tag=${1:-r02p}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${tag}_pytest_gpu.txt
for t in 1 0; do
  echo "== B200FE_MASS_KERNEL=$t" | tee -a gpurun_out/${tag}_mass_ab.txt
  B200FE_MASS_KERNEL=$t python tools/bk_bench.py --kinds bk1 --degrees 2,3,4,5,6,7 --reps 10 2>&1 | tail -n +2 | tee -a gpurun_out/${tag}_mass_ab.txt
  B200FE_MASS_KERNEL=$t python tools/op_sweep.py --degrees 2,3,4,5,6 --mass 2>&1 | grep "bp1" | tee -a gpurun_out/${tag}_mass_ab.txt
done
