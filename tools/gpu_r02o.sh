#!/bin/bash
# 1 GPU: A/B of 2 x 255 registers for the plain-contraction nq = 9, 10 Laplace kernels (E-vector BK3 p = 7, 8 with the drivers' cos() matrices)
tag=${1:-r02o}
mkdir -p gpurun_out
for v in default t2; do
  lib=""; [ $v != default ] && lib=$PWD/benchmarks_b200/variants/libb200fe_$v.so
  echo "== $v" | tee -a gpurun_out/${tag}_ab.txt
  B200FE_LIB=$lib python tools/bk_bench.py --kinds bk3 --degrees 6,7,8 --reps 10 2>&1 | tail -n +2 | tee -a gpurun_out/${tag}_ab.txt
done
