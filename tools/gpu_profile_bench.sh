#!/bin/bash
# ncu evidence for bench.py (1 GPU): launch list of the timed step + one full capture of the cell kernel.
# usage: bash tools/gpu_profile_bench.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --its 10 --no-sweep --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 200 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sumfact -s 40 -c 1 -f -o gpurun_out/${tag}_bp5_p6_kernel $CMD > gpurun_out/${tag}_full.log 2>&1
ls -la gpurun_out/${tag}_*
