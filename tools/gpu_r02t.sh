#!/bin/bash
# r02t: device-side index table + faster host numbering + exclusive interior stores.
#   1 GPU: bash tools/gpu_r02t.sh 1 r02t   -> parity suite, default bench line (setup_s)
#   N GPUs: bash tools/gpu_r02t.sh N r02t  -> dist_check (both transports), bench at N (parity block with the golden count, setup_s)
N=${1:-1}; tag=${2:-r02t}
mkdir -p gpurun_out
if [ "$N" = 1 ]; then
  python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/${tag}_pytest_gpu.txt
  B200FE_SETUP_TRACE=1 python bench.py --no-sweep > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err
  grep "b200fe setup" gpurun_out/${tag}_bench_1gpu.err | head -8
else
  TR="timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  $TR --master-port 29511 tools/dist_check.py > gpurun_out/${tag}_dist_check_${N}gpu.log 2>&1; grep "DIST_CHECK\|FAIL\|Error\|error" gpurun_out/${tag}_dist_check_${N}gpu.log | head -20
  B200FE_SETUP_TRACE=1 $TR --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${tag}_bench_${N}gpu.json 2> gpurun_out/${tag}_bench_${N}gpu.err
  grep "b200fe setup" gpurun_out/${tag}_bench_${N}gpu.err | grep "rank 0" | head -8
fi
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench_${N}gpu.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "parity")}, "e2e", d["e2e"]["value"], "setup_s", d["config"].get("setup_s"), d["config"].get("transport"),
      "kernel ms", d["roofline"]["avg_launch_ms"], "excl", d["roofline"].get("exclusive_interior_stores"), "comm_ab", d.get("comm_ab"))
PY
