#!/bin/bash
# GPU call J (1 GPU): parity after the Gram / TRSM changes, isolated kernel numbers, bench line, ncu of the new Gram kernel
# and of the TRSM update kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_j.log 2>&1; tail -4 gpurun_out/pytest_j.log
python scripts/profile_kernels.py > gpurun_out/kernels_j.log 2>&1; grep -E "gram_se|gram_mat|fill|copy|append|trsm|potrf" gpurun_out/kernels_j.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; echo "bench exit $?"
python -c "
import json; b=json.load(open('gpurun_out/bench_j.json')); print(b['value'], b['ms_per_step'], b['roofline']['frac'], b['e2e']['value'], b['e2e']['ms_per_step'], b['extras']['gram_gbs'], b['extras']['cfg1_whole_design'], b['cpu_baseline']['value'])"
CMD="python scripts/profile_kernels.py --once"
$CMD > gpurun_out/plain_j.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:gram_kernel -c 1 -o gpurun_out/prof_r02_gram_v2 $CMD > gpurun_out/ncu_j1.log 2>&1
ncu --set full --clock-control none -k regex:sub_ws -s 4 -c 1 -o gpurun_out/prof_r02_trsm_sub_ws $CMD > gpurun_out/ncu_j2.log 2>&1
ls -la gpurun_out/*.ncu-rep
