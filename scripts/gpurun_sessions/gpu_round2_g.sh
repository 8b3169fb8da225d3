#!/bin/bash
# GPU call G (1 GPU): isolated kernel timings against measured ceilings, then ONE ncu pass over Gram / TRSM / Cholesky kernels.
set -x
mkdir -p gpurun_out
python scripts/profile_kernels.py > gpurun_out/kernels_g.log 2>&1; tail -40 gpurun_out/kernels_g.log
python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_g.log 2>&1; tail -3 gpurun_out/pytest_g.log
CMD="python scripts/profile_kernels.py --once"
$CMD > gpurun_out/plain_g.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'gram_kernel|dmma_core|tri_solve|potrf_diag|append_row' -c 40 -o gpurun_out/prof_r02_gram_trsm $CMD > gpurun_out/ncu_g.log 2>&1
tail -3 gpurun_out/ncu_g.log
