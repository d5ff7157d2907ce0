#!/bin/bash
# 1 GPU: parity suite with the multi-component kernel, C5 key with / without it, full bench line
tag=${1:-r02g}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${tag}_pytest_gpu.txt
for mc in 1 0; do
B200FE_MULTI_COMPONENT=$mc python tools/bench_c5.py --cells-log2 5 --refine-frac 2 --its 50 --steps 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C5 1 GPU multi_component=$mc', {k: d[k] for k in d if k not in ('config',)})" | tee -a gpurun_out/${tag}_c5.txt
done
python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err
python -c "
import json
d = json.loads(open('gpurun_out/${tag}_bench_1gpu.json').read().strip().splitlines()[-1])
print('headline', d['value'], d['roofline']['frac'], d['e2e']['value']); print(d.get('bp6_hanging_nodes_p8'))
print('bk1', [round(r['frac_of_hbm_roofline'], 3) for r in d['degree_sweep_bk1_evector']])
print('ktb', [(r['kind'], r['p'], round(r['speedup'], 2)) for r in d['kernel_to_beat']['kernels'] if r['speedup'] < 1.1])
for r in d['degree_sweep_apply']: print(r['op'], r['p'], round(r['gdofs'], 2), round(r['frac_of_hbm_roofline'], 3))
"
