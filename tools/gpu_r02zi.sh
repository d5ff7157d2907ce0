#!/bin/bash
# 1 GPU: register-floor variants of the separable kernels (va: 80 registers, vb: 64) against the default build, on-the-fly sweep
tag=${1:-r02zi}
mkdir -p gpurun_out
python tools/op_sweep.py --geometry affine --json gpurun_out/${tag}_sweep_default.json > gpurun_out/${tag}_sweep_default.txt 2>&1
for v in va vb; do
  B200FE_LIB=$PWD/benchmarks_b200/variants/libb200fe_$v.so python tools/op_sweep.py --geometry affine --json gpurun_out/${tag}_sweep_$v.json > gpurun_out/${tag}_sweep_$v.txt 2>&1
done
paste <(cut -c1-72 gpurun_out/${tag}_sweep_default.txt) <(cut -c12-72 gpurun_out/${tag}_sweep_va.txt) <(cut -c12-72 gpurun_out/${tag}_sweep_vb.txt) | grep -v bp35 | tee gpurun_out/${tag}_sweep_table.txt
