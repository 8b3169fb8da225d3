#!/bin/bash
# GPU call F (8 GPUs): the N = 8 bench line (strong scaling + parity + extras.cfg5 north-star step) and cfg-4 at full size.
set -x
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_f_n8.json 2> gpurun_out/bench_f_n8.err
echo "bench exit $?"; tail -c 800 gpurun_out/bench_f_n8.err; python -c "
import json; b=json.load(open('gpurun_out/bench_f_n8.json')); print(b['value'], b['ms_per_step'], b['parity']); print(b['extras'].get('cfg5'))"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 scripts/cfg4_mi.py --V 200000 --N 512 > gpurun_out/cfg4_f.log 2>&1; tail -1 gpurun_out/cfg4_f.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 5 --warmup 3 --no-parity > gpurun_out/bench_f_n4.json 2> gpurun_out/bench_f_n4.err
python -c "
import json; b=json.load(open('gpurun_out/bench_f_n4.json')); print('n4', b['value'], b['ms_per_step'])"
