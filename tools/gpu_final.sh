#!/bin/bash
# round-end evidence on 1 GPU: full parity suite, smoke(), default bench line + reference arm.  usage: bash tools/gpu_final.sh <tag>
tag=${1:-r02r}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/${tag}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${tag}_smoke.txt
( time python bench.py > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err ) 2>&1 | grep real
( time python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err ) 2>&1 | grep real
python -c "
import json
d = json.loads(open('gpurun_out/${tag}_bench_1gpu.json').read().strip().splitlines()[-1])
print('headline', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], 'clocks', d['clocks'])
print('parity', d['parity']['its'], d['parity']['golden'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
print('c5', d['bp6_hanging_nodes_p8']['gdofs'], d['bp6_hanging_nodes_p8']['frac_of_hbm_roofline'])
r = json.loads(open('gpurun_out/${tag}_bench_reference.json').read().strip().splitlines()[-1]); print('reference arm', r['value'], r['cpu_baseline']['cores'])
"
