#!/bin/bash
# GPU call B (1 GPU): parity suite, A/B of the operand pipelines (ring 0 / 1 / 2), ncu capture of the tensor-map ring.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_b.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_b.log
tail -8 gpurun_out/pytest_b.log
for r in 0 1 2; do
  GPX_IVAR_RING=$r python scripts/ivar_sweep.py 2 100000 100000 63,255,1023 > gpurun_out/sweep_d2_ring$r.log 2>&1
  tail -3 gpurun_out/sweep_d2_ring$r.log
done
for r in 0 1 2; do
  GPX_IVAR_RING=$r python scripts/ivar_sweep.py 10 125000 100000 4095 > gpurun_out/sweep_d10_ring$r.log 2>&1; tail -1 gpurun_out/sweep_d10_ring$r.log
done
export GPX_IVAR_RING=1
CMD="python bench.py --steps 2 --warmup 3 --quick-design --no-cpu"
$CMD > gpurun_out/plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ivar_ws -s 2 -c 1 -o gpurun_out/prof_r02_tmap $CMD > gpurun_out/ncu_b.log 2>&1
tail -3 gpurun_out/ncu_b.log
