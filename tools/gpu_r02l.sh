#!/bin/bash
# 1 GPU: thread-per-element kernel with staged JxW (mass / Helmholtz): parity suite + A/B
tag=${1:-r02l}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${tag}_pytest_gpu.txt
for t in 1 0; do
  echo "== B200FE_TPE=$t" | tee -a gpurun_out/${tag}_tpe_ab.txt
  B200FE_TPE=$t python tools/op_sweep.py --degrees 1,2 --mass 2>&1 | tee -a gpurun_out/${tag}_tpe_ab.txt
  B200FE_TPE=$t python tools/bk_bench.py --kinds bk1 --degrees 1 --reps 10 2>&1 | tail -n +2 | tee -a gpurun_out/${tag}_tpe_ab.txt
done
