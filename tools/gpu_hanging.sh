#!/bin/bash
# hanging-node / vector-valued paths on one GPU: parity tests, then the bp6 driver at BASELINE-C5-like sizes.
# usage: bash tools/gpu_hanging.sh <tag> [full]
tag=${1:-h}; mkdir -p gpurun_out
if [ "$2" = "full" ]; then
  timeout 170 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/${tag}_pytest.log
else
  timeout 120 python -m pytest tests/test_zz_hanging_gpu.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/${tag}_pytest.log
fi
( timeout 40 ./benchmarks_b200/drivers/bp6 8 30000000 80000000 1 0.05; timeout 30 ./benchmarks_b200/drivers/bp6 8 30000000 80000000 0 0.05; timeout 30 ./benchmarks_b200/drivers/bp6 4 30000000 80000000 1 0.05 ) 2>&1 | tee gpurun_out/${tag}_bp6.log | tail -12
