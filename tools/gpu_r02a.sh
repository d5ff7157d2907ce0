#!/bin/bash
# round-2 first GPU job (1 GPU): parity suite, the reference's own CUDA kernels vs ours, tuning-variant A/B, DMMA ncu capture
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02a_pytest_gpu.txt
python tools/kernel_to_beat.py --json gpurun_out/r02a_kernel_to_beat.json 2>&1 | tee gpurun_out/r02a_kernel_to_beat.txt
bash tools/gpu_variants.sh r02a "eo eo_r112 eo_r136 nopad3" 2>&1 | tail -150
# one ncu capture each of the reference's DMMA kernel and of its CUDA-core twin
timeout 600 ncu --set full --clock-control none --import-source on -k regex:BwdTransHexKernel -c 4 -o gpurun_out/r02a_dmma \
    python -c "
import sys; sys.path.insert(0, '.')
import benchmarks_b200 as b
from oracle import ref_gpu
print(ref_gpu.dmma_study(b, ntests=1))
" > gpurun_out/r02a_dmma_ncu.log 2>&1
tail -3 gpurun_out/r02a_dmma_ncu.log
