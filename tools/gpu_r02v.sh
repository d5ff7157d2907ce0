#!/bin/bash
# 1 GPU: separable kernel for axis-aligned cells (QOP_CARTESIAN) -- parity suite, bench line (general affine vs cartesian keys), ncu capture
tag=${1:-r02v}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/${tag}_pytest_gpu.txt
python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err
python - <<PY | tee gpurun_out/${tag}_otf.txt
import json
d = json.loads(open("gpurun_out/${tag}_bench_1gpu.json").read().strip().splitlines()[-1])
print("headline %.2f e2e %.2f apply_only %s" % (d["value"], d["e2e"]["value"], d["apply_only"]))
o = d["apply_on_the_fly_affine_geometry"]
for k in ("gdofs", "ms", "speedup_vs_stored_G", "frac_of_hbm_roofline_own_bytes"): print("affine general", k, o.get(k))
print("cartesian", json.dumps(o.get("cartesian_cells"), indent=1))
print("trilinear", o.get("trilinear_cells", {}).get("gdofs"), o.get("trilinear_cells", {}).get("speedup_vs_stored_G"))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sumfact2 -s 6 -c 1 -f -o gpurun_out/${tag}_bp5_p6_cartesian python -c "
import sys; sys.path.insert(0, '.')
import torch, benchmarks_b200 as b
m = b.BoxMesh((1, 1, 1), 6, 6)
A = b.LaplaceOperator(m, quad='gll', with_jxw=False, geometry='affine')
src = torch.rand(m.n_owned, dtype=torch.float64, device='cuda'); dst = torch.empty_like(src)
for _ in range(8): A.vmult(dst, src)
torch.cuda.synchronize(); print(m.n_dofs_global, A.launch_info())
" > gpurun_out/${tag}_ncu_cartesian.log 2>&1; tail -2 gpurun_out/${tag}_ncu_cartesian.log
