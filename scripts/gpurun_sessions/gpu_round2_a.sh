#!/bin/bash
# GPU call A (1 GPU): parity suite, ring / prologue A-B sweep of the hot kernel, bench line, then ONE ncu capture.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_a.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_a.log
tail -5 gpurun_out/pytest_a.log
for r in 0 1 2; do
  GPX_IVAR_RING=$r python scripts/ivar_sweep.py 2 100000 100000 63,255,1023 > gpurun_out/sweep_d2_ring$r.log 2>&1
  tail -3 gpurun_out/sweep_d2_ring$r.log
done
GPX_IVAR_RING=0 GPX_FORCE_DIFF=1 python scripts/ivar_sweep.py 2 100000 100000 255 > gpurun_out/sweep_d2_ring0_diff.log 2>&1; tail -1 gpurun_out/sweep_d2_ring0_diff.log
for r in 0 2; do
  GPX_IVAR_RING=$r python scripts/ivar_sweep.py 10 125000 100000 4095 > gpurun_out/sweep_d10_ring$r.log 2>&1; tail -1 gpurun_out/sweep_d10_ring$r.log
done
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench exit $?"; tail -c 1500 gpurun_out/bench_a.json
CMD="python bench.py --steps 2 --warmup 3 --quick-design --no-cpu"
$CMD > gpurun_out/plain_a.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ivar_ws -s 2 -c 1 -o gpurun_out/prof_r02_base $CMD > gpurun_out/ncu_a.log 2>&1
tail -3 gpurun_out/ncu_a.log
