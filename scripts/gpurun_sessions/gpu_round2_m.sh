#!/bin/bash
# GPU call M (8 GPUs): final N = 8 and N = 4 bench lines.
bash scripts/gpurun_sessions/gpu_round2_l.sh 8
bash scripts/gpurun_sessions/gpu_round2_l.sh 4
