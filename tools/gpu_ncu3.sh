#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python benchmarks_b200/drivers/check_bk3.py 3 300000 2>&1 | tail -10
cap() { ncu --set full --clock-control none --import-source on -k regex:sumfact -s 3 -c 1 -f -o gpurun_out/$1 python tools/bk_bench.py --kinds $2 --degrees $3 --reps 2 > gpurun_out/$1.log 2>&1; }
cap ncu3_bk1_p3 bk1 3
cap ncu3_bk3_p4 bk3 4
cap ncu3_bk3_p8 bk3 8
ls gpurun_out/ncu3_*.ncu-rep
