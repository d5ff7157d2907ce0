#!/bin/bash
# ncu --set full captures of representative E-vector kernels (one launch each, after warm-up)
mkdir -p gpurun_out
cap() { # name lib kinds degrees
  local lib=""; [ "$2" != "default" ] && lib="B200FE_LIB=$PWD/benchmarks_b200/variants/libb200fe_$2.so"
  env $lib ncu --set full --clock-control none --import-source on -k regex:sumfact -s 3 -c 1 -f -o gpurun_out/$1 \
      python tools/bk_bench.py --kinds $3 --degrees $4 --reps 2 > gpurun_out/$1.log 2>&1
}
cap ncu_bk5_p6_default default bk5 6
cap ncu_bk5_p6_rollminb2 rollminb2 bk5 6
cap ncu_bk3_p8_rollminb2 rollminb2 bk3 8
cap ncu_bk1_p2_default default bk1 2
cap ncu_bk5_p3_tpb128 tpb128 bk5 3
ls -la gpurun_out/*.ncu-rep
