#!/bin/bash
# second hanging-node run: component-batched constraint kernels, deal.II-format geometry adapter, C5 bench pieces.
# usage: bash tools/gpu_hanging2.sh <tag>
tag=${1:-h2}; mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_zz_hanging_gpu.py "tests/test_operator_gpu.py::test_vector_valued_apply_bp6_style" -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/${tag}_pytest.log
timeout 25 ./benchmarks_b200/drivers/bp6 8 30000000 80000000 1 0.05 2>&1 | tee gpurun_out/${tag}_bp6.log | tail -3
timeout 30 python tools/bench_c5.py --cells-log2 5 --its 30 --steps 2 2>&1 | tee gpurun_out/${tag}_bench_c5_1gpu.json | tail -2
# launch list of one BP6 apply with hanging nodes (shares of distribute / cell / condense kernels)
timeout 45 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/${tag}_bp6_launches.csv \
    python tools/bench_c5.py --cells-log2 5 --its 2 --steps 1 --warmup 3 > gpurun_out/${tag}_bp6_launches.log 2>&1
timeout 50 python bench.py --no-cpu-baseline --cells-log2 4 --its 10 --steps 1 > gpurun_out/${tag}_bench_small.json 2> gpurun_out/${tag}_bench_small.err
python -c "
import json
try:
    print(json.dumps(json.load(open('gpurun_out/${tag}_bench_small.json')).get('bp6_hanging_nodes_p8')))
except Exception as e:
    print('bench small:', e)
"
