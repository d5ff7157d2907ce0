#!/bin/bash
# A/B of the tuning variants (built beforehand on the CPU box with tools/build_variants.sh): parity tests of the operator
# paths against the oracle with each variant library, then the operator-apply sweep.  usage: bash tools/gpu_variants.sh <tag> [variants]
tag=${1:-v}; variants=${2:-"eo nopad3 eo_r112 eo_r136"}
mkdir -p gpurun_out
python tools/op_sweep.py --json gpurun_out/${tag}_sweep_default.json | tee gpurun_out/${tag}_sweep_default.txt
for v in $variants; do
  lib=$PWD/benchmarks_b200/variants/libb200fe_$v.so
  [ -f "$lib" ] || { echo "missing $lib (run tools/build_variants.sh)"; continue; }
  echo "== variant $v"
  B200FE_LIB=$lib python -m pytest tests/test_operator_gpu.py tests/test_zz_hanging_gpu.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/${tag}_pytest_$v.log
  B200FE_LIB=$lib python tools/op_sweep.py --json gpurun_out/${tag}_sweep_$v.json | tee gpurun_out/${tag}_sweep_$v.txt
done
