#!/bin/bash
# GPU call H (1 GPU): Gram via the DMMA prologue vs the difference-form kernel; small ncu captures (one launch each).
set -x
mkdir -p gpurun_out
python scripts/profile_kernels.py > gpurun_out/kernels_h.log 2>&1; grep -E "gram|fill|copy" gpurun_out/kernels_h.log
CMD="python scripts/profile_kernels.py --once"
$CMD > gpurun_out/plain_h.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:gram_kernel -c 1 -o gpurun_out/prof_r02_gram $CMD > gpurun_out/ncu_h1.log 2>&1
ncu --set full --clock-control none -k regex:dmma_core -s 1 -c 2 -o gpurun_out/prof_r02_dmma_core $CMD > gpurun_out/ncu_h2.log 2>&1
ncu --set full --clock-control none -k regex:'tri_solve|potrf_diag' -s 8 -c 3 -o gpurun_out/prof_r02_trisolve_potrf $CMD > gpurun_out/ncu_h3.log 2>&1
ls -la gpurun_out/*.ncu-rep
