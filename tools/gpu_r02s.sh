#!/bin/bash
# 1 GPU: exclusive cell-interior stores -- parity suite, bench A/B (B200FE_EXCL_INTERIOR=1 / 0), ncu full capture of the headline kernel
tag=${1:-r02s}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/${tag}_pytest_gpu.txt
python bench.py > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err
B200FE_EXCL_INTERIOR=0 python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench_1gpu_atomics_only.json 2> gpurun_out/${tag}_bench_1gpu_atomics_only.err
python - <<PY | tee gpurun_out/${tag}_ab.txt
import json
for name in ("bench_1gpu", "bench_1gpu_atomics_only"):
    d = json.loads(open("gpurun_out/${tag}_%s.json" % name).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(name, "excl", r.get("exclusive_interior_stores"), "headline %.2f e2e %.2f" % (d["value"], d["e2e"]["value"]), "kernel ms %.4f frac %.3f share %.3f" % (r["avg_launch_ms"], r["frac"], r["kernel_share_of_step"]),
          "parity", d["parity"]["its"], d["parity"]["golden"], "C4 cg/apply", d["parity"].get("cg_gdofs"), d["parity"].get("apply_gdofs"), "clk", d["clocks"]["sm_mhz"])
    print("  c5", d["bp6_hanging_nodes_p8"]["gdofs"], d["bp6_hanging_nodes_p8"]["frac_of_hbm_roofline"])
    for k in ("degree_sweep_apply",):
        sw = d.get(k)
        if sw: print("  ", k, json.dumps(sw)[:1500])
PY
CMD="python bench.py --steps 2 --warmup 3 --its 10 --no-sweep --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:sumfact2 -s 40 -c 1 -f -o gpurun_out/${tag}_bp5_p6_kernel $CMD > gpurun_out/${tag}_full.log 2>&1
ls -la gpurun_out/${tag}_*
