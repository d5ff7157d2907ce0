#!/bin/bash
# hanging-node / vector-valued paths (BASELINE config C5) on one GPU: parity tests, the bp6 driver, the C5 bench tool with
# both constraint forms, and an ncu launch list of one constrained apply.  usage: bash tools/gpu_hanging.sh <tag> [full]
tag=${1:-h}; mkdir -p gpurun_out
if [ "$2" = "full" ]; then
  python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/${tag}_pytest.log
else
  python -m pytest tests/test_zz_hanging_gpu.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/${tag}_pytest.log
fi
( ./benchmarks_b200/drivers/bp6 8 30000000 80000000 1 0.05; ./benchmarks_b200/drivers/bp6 8 30000000 80000000 0 0.05 ) 2>&1 | tee gpurun_out/${tag}_bp6.log | tail -8
for form in faces rows; do
  python tools/bench_c5.py --cells-log2 5 --its 30 --steps 2 --constraints $form 2>&1 | tail -1 | tee gpurun_out/${tag}_bench_c5_1gpu_${form}.json
done
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/${tag}_bp6_launches.csv \
    python tools/bench_c5.py --cells-log2 5 --its 2 --steps 1 --warmup 3 > gpurun_out/${tag}_bp6_launches.log 2>&1
