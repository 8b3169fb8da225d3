#!/bin/bash
# GPU call D (2 GPUs): the strong-scaling bench line with the parity block (sharded engines vs oracle, native NCCL exchange).
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fitc" > gpurun_out/pytest_d.log 2>&1; tail -3 gpurun_out/pytest_d.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_d_n2.json 2> gpurun_out/bench_d_n2.err
echo "bench exit $?"; tail -c 1500 gpurun_out/bench_d_n2.err; tail -c 2500 gpurun_out/bench_d_n2.json
