#!/bin/bash
# 1 GPU: parity suite (incl. the C++ driver with Geometry::OnTheFly), config C5 on one GPU with the exclusive interior stores
tag=${1:-r02y}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/${tag}_pytest_gpu.txt
python tools/bench_c5.py --cells-log2 6 --refine-frac 4 --its 50 --steps 3 > gpurun_out/${tag}_bench_c5_1gpu.json 2> gpurun_out/${tag}_bench_c5_1gpu.err; head -c 900 gpurun_out/${tag}_bench_c5_1gpu.json; echo
B200FE_EXCL_INTERIOR=0 python tools/bench_c5.py --cells-log2 6 --refine-frac 4 --its 50 --steps 3 > gpurun_out/${tag}_bench_c5_1gpu_atomics_only.json 2> gpurun_out/${tag}_bench_c5_1gpu_atomics_only.err; head -c 300 gpurun_out/${tag}_bench_c5_1gpu_atomics_only.json; echo
B200FE_GEOMETRY=onthefly ./benchmarks_b200/drivers/bp3 4 1000000 20000000 1 gll 2>&1 | tail -12 | tee gpurun_out/${tag}_bp3_cxx_onthefly_gll.txt
./benchmarks_b200/drivers/bp3 4 1000000 20000000 1 gll 2>&1 | tail -12 | tee gpurun_out/${tag}_bp3_cxx_stored_gll.txt
