#!/bin/bash
# final-state check: GPU suite and smoke with the one-kernel loop on by default, then the default bench line
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 240 > gpurun_out/pytest_r.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_r.log; tail -4 gpurun_out/pytest_r.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke_r.log
timeout 900 python bench.py > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err; echo "bench exit $?"; tail -c 3000 gpurun_out/bench_r.json
