#!/bin/bash
# 1 GPU: separable kernel on the nodal values (interpolated operators on axis-aligned cells) -- parity suite, operator sweeps
# stored vs on-the-fly, register-floor / CTA-size variants of the new kernel
tag=${1:-r02z}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/${tag}_pytest_gpu.txt
python tools/op_sweep.py --json gpurun_out/${tag}_sweep_stored.json > gpurun_out/${tag}_sweep_stored.txt 2>&1
python tools/op_sweep.py --geometry affine --json gpurun_out/${tag}_sweep_cartesian.json > gpurun_out/${tag}_sweep_cartesian.txt 2>&1
for v in cart_r255 cart_r200 cart_t192; do
  B200FE_LIB=$PWD/benchmarks_b200/variants/libb200fe_$v.so python tools/op_sweep.py --geometry affine --json gpurun_out/${tag}_sweep_$v.json > gpurun_out/${tag}_sweep_$v.txt 2>&1
done
paste <(cut -c1-40 gpurun_out/${tag}_sweep_stored.txt) <(cut -c6-70 gpurun_out/${tag}_sweep_cartesian.txt) | tee gpurun_out/${tag}_sweep_table.txt
for v in cart_r255 cart_r200 cart_t192; do echo == $v; grep "bp3 \|bp35" gpurun_out/${tag}_sweep_$v.txt | cut -c1-75; done | tee -a gpurun_out/${tag}_sweep_table.txt
