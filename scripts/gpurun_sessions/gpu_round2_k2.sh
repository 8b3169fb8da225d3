#!/bin/bash
set -x
mkdir -p gpurun_out
SECONDS=0
python bench.py > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; echo "bench exit $? after ${SECONDS}s"; tail -5 gpurun_out/bench_k.err
python -c "
import json; b=json.load(open('gpurun_out/bench_k.json')); print(b['value'], b['ms_per_step'], b['roofline']['frac'], b['e2e']['value'], b['gpu_launches'], b['cpu_baseline']['value'], b['cpu_baseline']['kind']); print(b['extras']['cfg3']); print(b['extras']['cfg4']); print(b['extras']['cfg1_whole_design'])"
