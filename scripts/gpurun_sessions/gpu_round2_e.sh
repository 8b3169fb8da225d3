#!/bin/bash
# GPU call E (2 GPUs): block-cyclic MI engine -- unit tests on one GPU, |V| = 80 000 set-up on two (r01: 10.65 s), bench parity.
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "mi or dgemm or abi" > gpurun_out/pytest_e.log 2>&1; tail -4 gpurun_out/pytest_e.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/cfg4_mi.py --V 80000 --N 128 > gpurun_out/cfg4_e.log 2>&1; tail -2 gpurun_out/cfg4_e.log
GPX_MI_BLK=1024 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/cfg4_mi.py --V 80000 --N 128 > gpurun_out/cfg4_e_blk1024.log 2>&1; tail -1 gpurun_out/cfg4_e_blk1024.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 scripts/cfg4_mi.py --V 40000 --N 128 --compare-dense > gpurun_out/cfg4_e_dense.log 2>&1; tail -1 gpurun_out/cfg4_e_dense.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_e_n2.json 2> gpurun_out/bench_e_n2.err
echo "bench exit $?"; python -c "
import json; b=json.load(open('gpurun_out/bench_e_n2.json')); print(b['value'], b['parity'])"
