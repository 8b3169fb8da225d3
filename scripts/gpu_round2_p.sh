#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_p.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_p.log; tail -6 gpurun_out/pytest_p.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_p.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke_p.log
