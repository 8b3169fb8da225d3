#!/bin/bash
# GPU validation of the final state: the GPU test suite, smoke(), cfg-1 through both loops (each under its own timeout).
set -x
mkdir -p gpurun_out
GPX_TEST_ONE_KERNEL=1 timeout 600 python -m pytest tests -m gpu -q -x --timeout 240 > gpurun_out/pytest_p.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_p.log; tail -6 gpurun_out/pytest_p.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_p.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke_p.log
timeout 120 python scripts/cfg1_steps.py
