#!/bin/bash
# round-2 multi-GPU job.  usage: bash tools/gpu_r02b.sh <N> <tag> [quick]
N=${1:-2}; tag=${2:-r02b}; quick=${3:-}
mkdir -p gpurun_out
TR="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 tools/dist_check.py > gpurun_out/${tag}_dist_check_${N}gpu.log 2>&1; grep "DIST_CHECK\|FAIL\|Error\|error" gpurun_out/${tag}_dist_check_${N}gpu.log | head -20
$TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${tag}_bench_${N}gpu.json 2> gpurun_out/${tag}_bench_${N}gpu.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_${N}gpu.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "e2e", "parity", "comm_ab")}, d["config"].get("setup_s"), d["config"].get("transport"))
except Exception as e:
    print("bench line unreadable:", e); print(open("gpurun_out/${tag}_bench_${N}gpu.err").read()[-3000:])
PY
# p-halox: NCCL and P2P side by side, payload verified
KBS=1,16,128,1024,8192,65536,524288; [ -n "$quick" ] && KBS=1,16,128,1024,8192
PHALOX_SWEEP_DIMS=1,2,3 PHALOX_SWEEP_KB=$KBS timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 \
    benchmarks_b200/drivers/phalox.py 2 64 10 1 10 0 2>gpurun_out/${tag}_phalox_${N}gpu.err | grep "^P=" > gpurun_out/${tag}_phalox_${N}gpu.txt
grep "KB= 1 " gpurun_out/${tag}_phalox_${N}gpu.txt | cut -c1-400
grep -c "payload= ok" gpurun_out/${tag}_phalox_${N}gpu.txt; grep -c "WRONG" gpurun_out/${tag}_phalox_${N}gpu.txt
# the C++ driver on N ranks (reference protocol: bp3 <degree> <minsize> <maxsize>)
timeout 300 python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 \
    ./benchmarks_b200/drivers/bp3 4 1000000 40000000 > gpurun_out/${tag}_bp3_cxx_${N}gpu.log 2>&1; tail -6 gpurun_out/${tag}_bp3_cxx_${N}gpu.log
# config C5, strong scaling
$TR --master-port 29514 tools/bench_c5.py --cells-log2 6 --refine-frac 4 --its 50 --steps 3 > gpurun_out/${tag}_bench_c5_${N}gpu.json 2> gpurun_out/${tag}_bench_c5_${N}gpu.err; head -c 600 gpurun_out/${tag}_bench_c5_${N}gpu.json; echo; tail -3 gpurun_out/${tag}_bench_c5_${N}gpu.err
