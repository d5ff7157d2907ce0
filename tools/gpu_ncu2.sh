#!/bin/bash
mkdir -p gpurun_out
cap() { # name lib kinds degrees
  local lib=""; [ "$2" != "default" ] && lib="B200FE_LIB=$PWD/benchmarks_b200/variants/libb200fe_$2.so"
  env $lib ncu --set full --clock-control none --import-source on -k regex:sumfact -s 3 -c 1 -f -o gpurun_out/$1 \
      python tools/bk_bench.py --kinds $3 --degrees $4 --reps 2 > gpurun_out/$1.log 2>&1
}
cap ncu2_bk3_p3 tpb96 bk3 3
cap ncu2_bk3_p8 tpb96 bk3 8
cap ncu2_bk5_p6 tpb96 bk5 6
cap ncu2_bk1_p2 tpb96 bk1 2
for v in tpb64 tpb128 tpb96r96 tpb96r128; do
  B200FE_LIB=$PWD/benchmarks_b200/variants/libb200fe_$v.so python tools/bk_bench.py --json gpurun_out/bk2_$v.json > gpurun_out/bk2_$v.txt 2>&1
  B200FE_LIB=$PWD/benchmarks_b200/variants/libb200fe_$v.so python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench2_$v.json 2> gpurun_out/bench2_$v.err
done
ls gpurun_out/*.ncu-rep | tail -5
