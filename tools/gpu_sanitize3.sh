#!/bin/bash
# compute-sanitizer on the kernels added late in round 2: separable kernels for axis-aligned cells (collocated branch of
# sumfact2, nodal kernel sumfact_cart incl. mass / Helmholtz, cartesian diagonal / rhs), exclusive interior stores with the
# verification pass, device-side index-table expansion.  Small parity cases only.
mkdir -p gpurun_out
SEL='(separable and (mass-2 or mass-8 or helmholtz-5 or helmholtz-8)) or (sheared and (gauss-3 or gll-7 or gll-3 or gauss-8)) or (affine_geometry_matches and (gll-6 or gauss-2 or gauss-8 or gll-3)) or exclusive_interior or index_table'
for tool in memcheck racecheck; do
  extra=""; [ $tool = racecheck ] && extra="--racecheck-report all"
  timeout 600 compute-sanitizer --tool $tool $extra --print-limit 200 python -m pytest tests/test_operator_gpu.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitize3_${tool}_operator.log 2>&1; echo "$tool operator rc=$?"
  grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" gpurun_out/sanitize3_${tool}_operator.log | tail -3
done
