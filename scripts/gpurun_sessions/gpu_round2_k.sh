#!/bin/bash
# GPU call K (1 GPU): what the driver runs at round end -- the GPU test suite, smoke(), the default bench line.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_k.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_k.log; tail -4 gpurun_out/pytest_k.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_k.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_k.log
/usr/bin/time -v python bench.py > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; echo "bench exit $?"; grep -E "Elapsed|Maximum resident" gpurun_out/bench_k.err
python -c "
import json; b=json.load(open('gpurun_out/bench_k.json')); print(b['value'], b['ms_per_step'], b['roofline']['frac'], b['e2e']['value'], b['gpu_launches'], b['cpu_baseline']['value'], b['cpu_baseline']['kind']); print(b['extras']['cfg3']); print(b['extras']['cfg4']); print(b['extras']['cfg1_whole_design'])"
