#!/bin/bash
# GPU call L (N GPUs, N = $1): the bench line the driver's scaling run produces.
N=${1:-2}
set -x
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_l_n$N.json 2> gpurun_out/bench_l_n$N.err
echo "bench exit $? after ${SECONDS}s"; tail -3 gpurun_out/bench_l_n$N.err
python -c "
import json,sys; b=json.load(open('gpurun_out/bench_l_n$N.json')); print(b['value'], b['ms_per_step'], b['roofline']['frac'], b['e2e']['value'], b['gpu_launches']); print(b['parity']); print(b['extras'].get('cfg3')); print(b['extras'].get('cfg4')); print(b['extras'].get('cfg5'))"
