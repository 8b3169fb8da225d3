#!/bin/bash
# GPU call C (1 GPU): full parity suite, smoke(), the default bench line (with the reference arm as its cpu_baseline).
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_c.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_c.log
tail -15 gpurun_out/pytest_c.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_c.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke_c.log
python bench.py > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; echo "bench exit $?"; tail -c 600 gpurun_out/bench_c.err; tail -c 3000 gpurun_out/bench_c.json
