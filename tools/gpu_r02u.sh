#!/bin/bash
# 1 GPU: alternating pass directions in the CG loop (rejected experiment; the code it toggled was removed) -- kept as the record of the A/B
tag=${1:-r02u}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/${tag}_pytest_gpu.txt
for rep in 1 2; do
python bench.py --no-sweep --no-cpu-baseline --steps 10 > gpurun_out/${tag}_bench_alt_$rep.json 2> gpurun_out/${tag}_bench_alt_$rep.err
B200FE_CG_ALTERNATE=0 python bench.py --no-sweep --no-cpu-baseline --steps 10 > gpurun_out/${tag}_bench_fwd_$rep.json 2> gpurun_out/${tag}_bench_fwd_$rep.err
done
python - <<PY | tee gpurun_out/${tag}_ab.txt
import json
for name in ("alt_1", "fwd_1", "alt_2", "fwd_2"):
    d = json.loads(open("gpurun_out/${tag}_bench_%s.json" % name).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(name, "headline %.2f e2e %.2f" % (d["value"], d["e2e"]["value"]), "ms/step %.2f" % d["ms_per_step"], "kernel ms %.4f frac %.3f share %.3f" % (r["avg_launch_ms"], r["frac"], r["kernel_share_of_step"]), "clk", d["clocks"]["sm_mhz"])
PY
