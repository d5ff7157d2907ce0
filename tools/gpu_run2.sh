#!/bin/bash
# v2 kernel: parity tests + BK sweep + BP5 bench for tuning variants
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest.log
tail -4 gpurun_out/r2_pytest.log
python tools/bk_bench.py --json gpurun_out/bk2_default.json > gpurun_out/bk2_default.txt 2>&1
for v in "$@"; do
  B200FE_LIB=$PWD/benchmarks_b200/variants/libb200fe_$v.so python tools/bk_bench.py --json gpurun_out/bk2_$v.json > gpurun_out/bk2_$v.txt 2>&1
done
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench2_default.json 2> gpurun_out/bench2_default.err
for v in "$@"; do
  B200FE_LIB=$PWD/benchmarks_b200/variants/libb200fe_$v.so python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench2_$v.json 2> gpurun_out/bench2_$v.err
done
tail -3 gpurun_out/bench2_default.err
