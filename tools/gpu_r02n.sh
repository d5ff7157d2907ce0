#!/bin/bash
# 1 GPU: p-multigrid transfer + V-cycle parity, then PMG / Chebyshev / Jacobi / identity CG at size (BP5 p = 4 and 8)
tag=${1:-r02n}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${tag}_pytest_gpu.txt
python - <<PY 2>&1 | tee gpurun_out/${tag}_preconditioners.txt
import sys, time; sys.path.insert(0, '.')
import torch, benchmarks_b200 as b
for degrees, cyc in (((4, 2, 1), 15), ((8, 4, 2, 1), 12)):
    ops = [b.LaplaceOperator(b.BoxMesh.bp3_cycle(cyc, p), quad='gll') for p in degrees]
    A = ops[0]; rhs = A.compute_rhs()
    def run(name, pre):
        x = A.initialize_dof_vector(); ctl = b.ReductionControl(20000, 1e-16, 1e-9)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        b.SolverCG(ctl).solve(A, x, rhs, pre)
        torch.cuda.synchronize(); t = time.perf_counter() - t0
        print(f'BP5 p={degrees[0]} {A.mesh.n_dofs_global} DoFs  {name:34s} its {ctl.last_step():5d}  time {t:8.4f} s', flush=True)
    run('identity', None)
    run('Jacobi', A.get_matrix_diagonal_inverse())
    run('Chebyshev(5)', b.PreconditionChebyshev(A, degree=5, smoothing_range=20.0))
    run(f'p-multigrid {degrees}, Chebyshev(3)', b.PreconditionPMG(ops, smoother_degree=3, smoothing_range=20.0, coarse_degree=10))
PY
