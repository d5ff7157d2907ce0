#!/bin/bash
# multi-GPU evidence: parity check, scaling bench, p-halox sweep.  usage: bash tools/gpu_multi.sh <ngpus> <tag>
N=${1:-2}; tag=${2:-r01}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 tools/dist_check.py > gpurun_out/${tag}_dist_check_${N}gpu.log 2>&1; tail -2 gpurun_out/${tag}_dist_check_${N}gpu.log
# the C++ driver on N ranks (reference protocol: bp3 <degree> <minsize> <maxsize>)
timeout 300 python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 \
    ./benchmarks_b200/drivers/bp3 4 1000000 40000000 > gpurun_out/${tag}_bp3_cxx_${N}gpu.log 2>&1; tail -4 gpurun_out/${tag}_bp3_cxx_${N}gpu.log
# config C5, strong scaling
$TR --master-port 29514 tools/bench_c5.py --cells-log2 6 --refine-frac 4 --its 50 --steps 3 > gpurun_out/${tag}_bench_c5_${N}gpu.json 2> gpurun_out/${tag}_bench_c5_${N}gpu.err; head -c 300 gpurun_out/${tag}_bench_c5_${N}gpu.json; echo
$TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${tag}_bench_${N}gpu.json 2> gpurun_out/${tag}_bench_${N}gpu.err
head -c 300 gpurun_out/${tag}_bench_${N}gpu.json; echo
$TR --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 --overlap 0 > gpurun_out/${tag}_bench_${N}gpu_nooverlap.json 2> gpurun_out/${tag}_bench_${N}gpu_nooverlap.err
head -c 200 gpurun_out/${tag}_bench_${N}gpu_nooverlap.json; echo
: > gpurun_out/${tag}_phalox_${N}gpu.txt
port=29520
for dim in 1 2 3; do for kb in 1 16 128 1024 8192 65536; do
  $TR --master-port $port benchmarks_b200/drivers/phalox.py $dim $kb 10 1 10 0 2>/dev/null | grep "^P=" >> gpurun_out/${tag}_phalox_${N}gpu.txt; port=$((port+1))
done; done
tail -4 gpurun_out/${tag}_phalox_${N}gpu.txt
