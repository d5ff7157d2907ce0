#!/bin/bash
# 1 GPU: C5 at the strong-scaling mesh, final bench line + reference arm + ncu launch list and full capture, ncu of the p = 1 kernels
tag=${1:-r02j}
mkdir -p gpurun_out
python tools/bench_c5.py --cells-log2 6 --refine-frac 4 --its 50 --steps 3 > gpurun_out/${tag}_bench_c5_1gpu.json 2> gpurun_out/${tag}_bench_c5_1gpu.err; head -c 700 gpurun_out/${tag}_bench_c5_1gpu.json; echo
python bench.py > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err
python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
python -c "
import json
d = json.loads(open('gpurun_out/${tag}_bench_1gpu.json').read().strip().splitlines()[-1])
print('headline', d['value'], d['roofline']['frac'], d['e2e']['value'], d['config']['setup_s']); print(d.get('bp6_hanging_nodes_p8')); print(d.get('apply_on_the_fly_affine_geometry'))
r = json.loads(open('gpurun_out/${tag}_bench_reference.json').read().strip().splitlines()[-1]); print('reference arm', r['value'], r['cpu_baseline']['cores'])
"
CMD="python bench.py --steps 2 --warmup 3 --its 10 --no-sweep --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sumfact2 -s 40 -c 1 -f -o gpurun_out/${tag}_bp5_p6_kernel $CMD > gpurun_out/${tag}_full.log 2>&1
# p = 1: what binds the tiny-element kernels (E-vector BK3 and L-vector BP5)
timeout 300 ncu --set full --clock-control none -k regex:sumfact2 -s 3 -c 1 -f -o gpurun_out/${tag}_bk3_p1 python -c "
import sys; sys.path.insert(0, '.')
import numpy as np, torch, benchmarks_b200 as b
p, nq = 1, 3; nelmt = 1250000
basis = np.cos(np.arange(nq * 2, dtype=np.float64)); dbasis = np.cos(np.arange(nq * nq, dtype=np.float64))
u = torch.rand(nelmt * 8, dtype=torch.float64, device='cuda'); G = torch.rand(nelmt * 6 * 27, dtype=torch.float64, device='cuda'); o = torch.empty_like(u)
for _ in range(6): b.bk3_apply(p, nq, basis, dbasis, G, u, o)
torch.cuda.synchronize()
" > gpurun_out/${tag}_ncu_bk3_p1.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:sumfact2 -s 3 -c 1 -f -o gpurun_out/${tag}_bp5_p1 python -c "
import sys; sys.path.insert(0, '.')
import torch, benchmarks_b200 as b
m = b.BoxMesh.bp3_cycle(23, 1)
A = b.LaplaceOperator(m, quad='gll', with_jxw=False)
src = torch.rand(m.n_owned, dtype=torch.float64, device='cuda'); dst = torch.empty_like(src)
for _ in range(6): A.vmult(dst, src)
torch.cuda.synchronize(); print(m.n_dofs_global, A.launch_info())
" > gpurun_out/${tag}_ncu_bp5_p1.log 2>&1; tail -1 gpurun_out/${tag}_ncu_bp5_p1.log
ls -la gpurun_out/${tag}_*
