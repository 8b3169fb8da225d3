#!/bin/bash
# one-kernel greedy loop: parity test against the multi-launch loop, then cfg-1 timing through all three loops
set -x
mkdir -p gpurun_out
GPX_TEST_ONE_KERNEL=1 timeout 300 python -m pytest tests -m gpu -q -x --timeout 120 -k "one_kernel or c_side_loop" > gpurun_out/pytest_q.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_q.log; tail -4 gpurun_out/pytest_q.log
timeout 120 python scripts/cfg1_steps.py 2>&1 | tee gpurun_out/cfg1_q.log
