#!/bin/bash
# round-2 1-GPU job: the full bench line (new keys), reference arm, ncu captures of the even-odd kernels at p = 7, 8
tag=${1:-r02c}
mkdir -p gpurun_out
( time python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err ) 2>&1 | tail -3
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_1gpu.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "parity")}); print(d["roofline"]["frac"], d["e2e"]["value"], d.get("cpu_baseline"))
    for r in d.get("degree_sweep_apply", []): print(r["op"], r["p"], round(r["gdofs"], 2), round(r["frac_of_hbm_roofline"], 3), r.get("cg_gdofs"))
    for k in ("degree_sweep_bk1_evector", "degree_sweep_bk3_evector", "degree_sweep_bk5_evector", "degree_sweep_bk1_evector_real_basis", "degree_sweep_bk3_evector_real_basis"):
        print(k, [round(r["frac_of_hbm_roofline"], 3) for r in d.get(k, [])])
    print("ktb", [(r["kind"], r["p"], round(r["speedup"], 2)) for r in d.get("kernel_to_beat", {}).get("kernels", [])])
    print(d.get("apply_on_the_fly_affine_geometry")); print(d.get("bp6_hanging_nodes_p8"))
except Exception as e:
    print("bench line unreadable:", e); print(open("gpurun_out/${tag}_bench_1gpu.err").read()[-3000:])
PY
( time python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err ) 2>&1 | tail -3
head -c 600 gpurun_out/${tag}_bench_reference.json; echo
# ncu: BP3 p = 7 and p = 8 (even-odd kernels), one full capture each
for p in 7 8; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sumfact2 -s 3 -c 1 -f -o gpurun_out/${tag}_bp3_p${p}_eo python -c "
import sys; sys.path.insert(0, '.')
import torch, benchmarks_b200 as b
m = b.BoxMesh.bp3_cycle(${p} == 7 and 15 or 15, ${p})
A = b.LaplaceOperator(m, nq=${p} + 2, with_jxw=False)
src = torch.rand(m.n_owned, dtype=torch.float64, device='cuda'); dst = torch.empty_like(src)
for _ in range(6): A.vmult(dst, src)
torch.cuda.synchronize(); print(A.launch_info())
" > gpurun_out/${tag}_ncu_p${p}.log 2>&1; tail -2 gpurun_out/${tag}_ncu_p${p}.log
done
