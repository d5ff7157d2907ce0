#!/bin/bash
# 1 GPU: trilinear on-the-fly geometry (parity + bench key), reference-CUDA-kernel parity test, full bench line
tag=${1:-r02m}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${tag}_pytest_gpu.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err
python -c "
import json
d = json.loads(open('gpurun_out/${tag}_bench_1gpu.json').read().strip().splitlines()[-1])
print('headline', d['value'], d['roofline']['frac'], d['e2e']['value']); print(d.get('apply_on_the_fly_affine_geometry'))
for r in d['degree_sweep_apply']: print(r['op'], r['p'], round(r['gdofs'], 2), round(r['frac_of_hbm_roofline'], 3), r.get('cg_gdofs'))
for k in ('degree_sweep_bk1_evector', 'degree_sweep_bk3_evector', 'degree_sweep_bk5_evector', 'degree_sweep_bk1_evector_real_basis', 'degree_sweep_bk3_evector_real_basis'):
    print(k, [round(r['frac_of_hbm_roofline'], 3) for r in d.get(k, [])])
print('ktb', [(r['kind'], r['p'], round(r['speedup'], 2)) for r in d['kernel_to_beat']['kernels']])
" || tail -5 gpurun_out/${tag}_bench_1gpu.err
