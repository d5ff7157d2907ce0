#!/bin/bash
# round-2 multi-GPU evidence job.  usage: bash tools/gpu_r02h.sh <N> <tag> [phalox: full|quick|none]
N=${1:-8}; tag=${2:-r02h}; ph=${3:-full}
mkdir -p gpurun_out
TR="timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 tools/dist_check.py > gpurun_out/${tag}_dist_check_${N}gpu.log 2>&1; grep "DIST_CHECK\|FAIL\|Error\|error" gpurun_out/${tag}_dist_check_${N}gpu.log | head -20
$TR --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${tag}_bench_${N}gpu.json 2> gpurun_out/${tag}_bench_${N}gpu.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_${N}gpu.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "e2e", "parity", "comm_ab")}, d["config"].get("setup_s"), d["config"].get("transport"), d["roofline"]["avg_launch_ms"], d["roofline"]["launches_per_apply"])
except Exception as e:
    print("bench line unreadable:", e); print(open("gpurun_out/${tag}_bench_${N}gpu.err").read()[-3000:])
PY
# config C5 (BP6 p = 8, hanging nodes, ~225 M DoFs), strong scaling
$TR --master-port 29514 tools/bench_c5.py --cells-log2 6 --refine-frac 4 --its 50 --steps 3 > gpurun_out/${tag}_bench_c5_${N}gpu.json 2> gpurun_out/${tag}_bench_c5_${N}gpu.err; head -c 900 gpurun_out/${tag}_bench_c5_${N}gpu.json; echo; tail -2 gpurun_out/${tag}_bench_c5_${N}gpu.err
# the C++ driver on N ranks (reference protocol: bp3 <degree> <minsize> <maxsize>); up to the 135 M-DoF golden row at N = 8
MAXS=$((17500000 * N))
timeout 600 python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 \
    ./benchmarks_b200/drivers/bp3 4 8000000 $MAXS > gpurun_out/${tag}_bp3_cxx_${N}gpu.log 2>&1; grep -A3 "cg_its\|mv_ghost" gpurun_out/${tag}_bp3_cxx_${N}gpu.log | tail -24
if [ "$ph" != none ]; then
KBS=1,16,128,1024,8192,65536,524288; [ "$ph" = quick ] && KBS=1,16,128,1024,8192
PHALOX_SWEEP_DIMS=1,2,3 PHALOX_SWEEP_KB=$KBS timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 \
    benchmarks_b200/drivers/phalox.py 2 64 10 1 10 0 2>gpurun_out/${tag}_phalox_${N}gpu.err | grep "^P=" > gpurun_out/${tag}_phalox_${N}gpu.txt
python - <<PY
rows=[]
for l in open('gpurun_out/${tag}_phalox_${N}gpu.txt'):
    t=l.split(); d=dict(zip(t[0::2],t[1::2]))
    rows.append((int(d['dim=']),int(d['KB=']),d['mode='],d['transport='],float(d['max_time_s=']),float(d['agg_BW_GBps=']),d['payload=']))
print("dim KB mode | nccl us/round  p2p us/round | nccl GB/s p2p GB/s | payload")
for dim in (1,2,3):
    for kb in sorted({r[1] for r in rows}):
        for mode in ("stream",):
            a=[r for r in rows if r[:3]==(dim,kb,mode) and r[3]=='nccl']; b=[r for r in rows if r[:3]==(dim,kb,mode) and r[3]=='p2p']
            if a and b: print(dim,kb,mode,"| %.1f  %.1f | %.1f %.1f | %s %s"%(a[0][4]*1e5,b[0][4]*1e5,a[0][5],b[0][5],a[0][6],b[0][6]))
PY
fi
