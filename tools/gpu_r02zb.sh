#!/bin/bash
# 1 GPU: nodal separable kernel with packed even-odd matrices -- parity suite, operator sweep stored vs on-the-fly
tag=${1:-r02zb}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/${tag}_pytest_gpu.txt
python tools/op_sweep.py --json gpurun_out/${tag}_sweep_stored.json > gpurun_out/${tag}_sweep_stored.txt 2>&1
python tools/op_sweep.py --geometry affine --json gpurun_out/${tag}_sweep_cartesian.json > gpurun_out/${tag}_sweep_cartesian.txt 2>&1
paste <(cut -c1-40 gpurun_out/${tag}_sweep_stored.txt) <(cut -c6-70 gpurun_out/${tag}_sweep_cartesian.txt) | tee gpurun_out/${tag}_sweep_table.txt
