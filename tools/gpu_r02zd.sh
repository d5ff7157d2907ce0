#!/bin/bash
# 1 GPU: separable mass / Helmholtz operators -- parity suite, operator sweep (incl. BP1 and Helmholtz) stored vs on-the-fly, bp5 driver rows
tag=${1:-r02zd}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/${tag}_pytest_gpu.txt
python tools/op_sweep.py --mass --json gpurun_out/${tag}_sweep_stored.json > gpurun_out/${tag}_sweep_stored.txt 2>&1
python tools/op_sweep.py --mass --geometry affine --json gpurun_out/${tag}_sweep_cartesian.json > gpurun_out/${tag}_sweep_cartesian.txt 2>&1
paste <(cut -c1-40 gpurun_out/${tag}_sweep_stored.txt) <(cut -c6-70 gpurun_out/${tag}_sweep_cartesian.txt) | grep "bp1\|helm" | tee gpurun_out/${tag}_sweep_table.txt
for s in 30 36; do ./benchmarks_b200/drivers/bp5 4 $s 1 | tail -1; B200FE_GEOMETRY=onthefly ./benchmarks_b200/drivers/bp5 4 $s 1 | tail -1; done | tee gpurun_out/${tag}_bp5_cxx.txt
