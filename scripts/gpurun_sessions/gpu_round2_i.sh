#!/bin/bash
# GPU call I (1 GPU): parity after the Gram rewrite, isolated kernel numbers, then compute-sanitizer racecheck on the small case.
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/pytest_i.log 2>&1; tail -3 gpurun_out/pytest_i.log
python scripts/profile_kernels.py > gpurun_out/kernels_i.log 2>&1; grep -E "gram|fill|copy|append" gpurun_out/kernels_i.log
python scripts/sanitize_case.py > gpurun_out/sanitize_plain.log 2>&1 || { tail -20 gpurun_out/sanitize_plain.log; exit 1; }
tail -2 gpurun_out/sanitize_plain.log
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 50 python scripts/sanitize_case.py > gpurun_out/racecheck_r02.log 2>&1
echo "racecheck exit $?"; tail -15 gpurun_out/racecheck_r02.log
