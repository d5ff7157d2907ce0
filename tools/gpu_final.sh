#!/bin/bash
# round-end evidence on 1 GPU: full parity suite, smoke(), default bench line + reference arm, ncu launch list of the bench command,
# ncu --set full captures of the headline cell kernel and of the nodal separable kernel.  usage: bash tools/gpu_final.sh <tag>
tag=${1:-r02zz}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/${tag}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${tag}_smoke.txt
( time python bench.py > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err ) 2>&1 | grep real
( time python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err ) 2>&1 | grep real
python -c "
import json
d = json.loads(open('gpurun_out/${tag}_bench_1gpu.json').read().strip().splitlines()[-1])
print('headline', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], 'clocks', d['clocks'], 'setup_s', d['config']['setup_s'])
print('parity', d['parity']['its'], d['parity']['golden'], d['parity'].get('on_the_fly_cartesian'), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
print('c5', d['bp6_hanging_nodes_p8']['gdofs'], d['bp6_hanging_nodes_p8']['frac_of_hbm_roofline'])
print('otf', {k: v for k, v in d['apply_on_the_fly_affine_geometry'].get('cartesian_cells', {}).items() if k != 'what'})
r = json.loads(open('gpurun_out/${tag}_bench_reference.json').read().strip().splitlines()[-1]); print('reference arm', r['value'], r['cpu_baseline']['cores'])
"
CMD="python bench.py --steps 2 --warmup 3 --its 10 --no-sweep --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sumfact2 -s 40 -c 1 -f -o gpurun_out/${tag}_bp5_p6_kernel $CMD > gpurun_out/${tag}_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sumfact_cart -s 6 -c 1 -f -o gpurun_out/${tag}_bp3_p4_cart python -c "
import sys; sys.path.insert(0, '.')
import torch, benchmarks_b200 as b
m = b.BoxMesh((1, 1, 1), 6, 4)
A = b.LaplaceOperator(m, nq=6, quad='gauss', with_jxw=False, geometry='affine')
src = torch.rand(m.n_owned, dtype=torch.float64, device='cuda'); dst = torch.empty_like(src)
for _ in range(8): A.vmult(dst, src)
torch.cuda.synchronize(); print(m.n_dofs_global, A.launch_info())
" > gpurun_out/${tag}_ncu_cart.log 2>&1; tail -1 gpurun_out/${tag}_ncu_cart.log
ls -la gpurun_out/${tag}_* | awk '{print $5, $9}'
