#!/bin/bash
# tests + low-degree sweep (E-vector) + operator apply sweep via bench (no CG baseline)
tag=${1:-q2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/bk_bench.py --kinds bk1,bk3,bk5 --degrees ${2:-1,2,3,4} --reps 10 --json gpurun_out/bk_$tag.json | tail -n +2
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$tag.json'))
sw=d['degree_sweep_apply']
print(f"value={d['value']:.2f} kernel_ms={d['roofline']['avg_launch_ms']:.3f} frac={d['roofline']['frac']:.3f}")
print("bp5:", " ".join(f"{r['frac_of_hbm_roofline']:.2f}" for r in sw if r['op']=='bp5'), "| bp3:", " ".join(f"{r['frac_of_hbm_roofline']:.2f}" for r in sw if r['op']=='bp3'))
PY
