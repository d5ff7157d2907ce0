#!/bin/bash
# compute-sanitizer on small parity cases: memcheck + racecheck (shared-memory hazards of the hand-placed barriers)
mkdir -p gpurun_out
SEL='(random_inputs and 37) or (vmult_matches and (bp3 or bp5 or helmholtz) and (2- or 5- or 7-)) or tail_batches'
timeout 280 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_bk_gpu.py tests/test_operator_gpu.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/sanitize_memcheck.log
timeout 280 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python -m pytest tests/test_bk_gpu.py tests/test_operator_gpu.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|hazard|passed|failed" gpurun_out/sanitize_racecheck.log | tail -5
