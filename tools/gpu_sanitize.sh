#!/bin/bash
# compute-sanitizer racecheck on small parity cases (shared-memory hazards of the hand-placed barriers)
mkdir -p gpurun_out
SEL='(random_inputs and 37) or (vmult_matches and (bp3 or bp5 or helmholtz or bp1) and (2- or 3- or 5- or 7-)) or tail_batches'
timeout 400 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 2000 python -m pytest tests/test_bk_gpu.py tests/test_operator_gpu.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_racecheck.log | tail -3
grep -E "(Write|Read) Thread" gpurun_out/sanitize_racecheck.log | sed -E 's/Thread \([0-9]+,0,0\)/Thread/; s/\+0x[0-9a-f]+//; s/\(b200fe::Mats.*//' | sort | uniq -c | sort -rn | head -20
