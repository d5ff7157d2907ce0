#!/bin/bash
# compute-sanitizer on the round-2 kernels: even-odd contractions (default for every operator built from a real basis), the
# multi-component kernel, diag(C^T A C), Chebyshev / deferred-x CG.  Small parity cases only (the tools slow kernels 10-100x).
mkdir -p gpurun_out
SEL='(vmult_matches and (bp3 or bp5 or helmholtz or bp1) and (2- or 4- or 7- or 8-)) or chebyshev or tail_batches or (golden and (0- or 1-))'
SELH='vector_valued or diagonal_with_constraints or (constrained_vmult and (bp5 or bp3) and (2- or 8-))'
for tool in memcheck racecheck; do
  extra=""; [ $tool = racecheck ] && extra="--racecheck-report all"
  timeout 700 compute-sanitizer --tool $tool $extra --print-limit 200 python -m pytest tests/test_operator_gpu.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitize2_${tool}_operator.log 2>&1; echo "$tool operator rc=$?"
  grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" gpurun_out/sanitize2_${tool}_operator.log | tail -3
  timeout 700 compute-sanitizer --tool $tool $extra --print-limit 200 python -m pytest tests/test_zz_hanging_gpu.py -m gpu -x -q -k "$SELH" > gpurun_out/sanitize2_${tool}_hanging.log 2>&1; echo "$tool hanging rc=$?"
  grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" gpurun_out/sanitize2_${tool}_hanging.log | tail -3
done
