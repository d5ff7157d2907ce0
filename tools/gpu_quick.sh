#!/bin/bash
# short confirmation run: GPU tests + a reduced sweep.  usage: bash tools/gpu_quick.sh <tag> [kinds] [degrees]
tag=${1:-q}; kinds=${2:-bk1,bk3}; degs=${3:-2,3,4,6,8}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/bk_bench.py --kinds $kinds --degrees $degs --reps 10 --json gpurun_out/bk_$tag.json | tail -n +2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
