#!/bin/bash
# 2 GPUs: p-halox with the in-launch multi-round exchange (mode = launch), small messages
N=${1:-2}; tag=${2:-r02ze}
mkdir -p gpurun_out
PHALOX_SWEEP_DIMS=1,2,3 PHALOX_SWEEP_KB=1,4,16,128 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 \
    benchmarks_b200/drivers/phalox.py 2 64 10 1 10 0 2>gpurun_out/${tag}_phalox_${N}gpu.err | grep "^P=" > gpurun_out/${tag}_phalox_${N}gpu.txt
tail -3 gpurun_out/${tag}_phalox_${N}gpu.err
python - <<PY
for l in open('gpurun_out/${tag}_phalox_${N}gpu.txt'):
    t=l.split(); d=dict(zip(t[0::2],t[1::2]))
    if d['mode=']!='sync': print(d['dim='],d['KB='],d['mode='],d['transport='],"%.2f us/round"%(float(d['max_time_s='])*1e6/int(d['nMsg='])),d.get('payload='))
PY
