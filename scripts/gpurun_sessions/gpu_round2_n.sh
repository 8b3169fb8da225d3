#!/bin/bash
# GPU call N (1 GPU): exact IVAR grid, narrow append blocks, more row segments -- parity, isolated numbers, bench.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_n.log 2>&1; tail -3 gpurun_out/pytest_n.log
python scripts/profile_kernels.py > gpurun_out/kernels_n.log 2>&1; grep -E "append|gram_se_2d_gbs" gpurun_out/kernels_n.log
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_n.json 2> gpurun_out/bench_n.err; echo "bench exit $?"
python -c "
import json; b=json.load(open('gpurun_out/bench_n.json')); print(b['value'], b['ms_per_step'], b['roofline']['frac'], b['e2e']['value'], b['extras']['append_row_gbs'], b['extras']['cfg1_whole_design'], b['extras']['resident_covariance_mode']['ms_per_step'], b['extras']['cfg3']['design_s'])"
