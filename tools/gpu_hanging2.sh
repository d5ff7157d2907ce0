#!/bin/bash
# second hanging-node run: batched constraint kernels.  usage: bash tools/gpu_hanging2.sh <tag>
tag=${1:-h2}; mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_zz_hanging_gpu.py "tests/test_operator_gpu.py::test_vector_valued_apply_bp6_style" -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/${tag}_pytest.log
timeout 25 ./benchmarks_b200/drivers/bp6 8 30000000 80000000 1 0.05 2>&1 | tee gpurun_out/${tag}_bp6.log | tail -3
timeout 50 python bench.py --no-cpu-baseline --cells-log2 4 --its 10 --steps 1 > gpurun_out/${tag}_bench_small.json 2> gpurun_out/${tag}_bench_small.err
python - <<PY
import json,sys
try:
    d = json.load(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/TAG_bench_small.json".replace("TAG", "${tag}")))
    print(json.dumps(d.get("bp6_hanging_nodes_p8")))
except Exception as e:
    print("bench small:", e)
PY
