#!/bin/bash
# round-end evidence on 1 GPU: bench line + reference arm + ncu launch list + one full capture of the cell kernel
tag=${1:-r01c}
mkdir -p gpurun_out
python bench.py > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err
python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
CMD="python bench.py --steps 2 --warmup 3 --its 10 --no-sweep --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 200 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sumfact -s 40 -c 1 -f -o gpurun_out/${tag}_bp5_p6_kernel $CMD > gpurun_out/${tag}_full.log 2>&1
head -c 400 gpurun_out/${tag}_bench_1gpu.json; echo; cat gpurun_out/${tag}_bench_reference.json | head -c 300
