#!/bin/bash
# first GPU exploration: parity tests + BK sweep for several launch-bound / unroll variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r1_pytest.log
cat gpurun_out/r1_pytest.log | tail -5
python tools/bk_bench.py --json gpurun_out/bk_default.json > gpurun_out/bk_default.txt 2>&1
for v in minb2 roll rollminb2 tpb128 tpb512 minb3 tpb128minb4; do
  B200FE_LIB=$PWD/benchmarks_b200/variants/libb200fe_$v.so python tools/bk_bench.py --json gpurun_out/bk_$v.json > gpurun_out/bk_$v.txt 2>&1
done
tail -30 gpurun_out/bk_default.txt
