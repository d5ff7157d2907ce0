#!/bin/bash
# 1 GPU: full parity suite (new: diag(C^T A C), Chebyshev), register-floor variant t1 vs default on the affected kernels
tag=${1:-r02e}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${tag}_pytest_gpu.txt
for v in default t1; do
  lib=""; [ $v != default ] && lib=$PWD/benchmarks_b200/variants/libb200fe_$v.so
  echo "== $v"
  B200FE_LIB=$lib python tools/op_sweep.py --degrees 8 --json gpurun_out/${tag}_sweep_$v.json | tee gpurun_out/${tag}_sweep_$v.txt
  B200FE_LIB=$lib python tools/bk_bench.py --kinds bk1 --degrees 6,7,8 --reps 10 --json gpurun_out/${tag}_bk_$v.json | tail -n +2 | tee -a gpurun_out/${tag}_sweep_$v.txt
done
python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu-baseline | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('headline', d['value'], 'ms/step', d['ms_per_step'], 'kernel share', d['roofline']['kernel_share_of_step'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], 'e2e', d['e2e']['value'])
" | tee gpurun_out/${tag}_headline.txt
