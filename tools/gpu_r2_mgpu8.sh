#!/bin/bash
# 8-GPU session (run under gpurun --gpus 8): multi == single parity over peer memory (both precisions), bench lines at 512^3 and 1024^3
N=8; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
{
timeout 200 $TR tools/mgpu_check.py 128 3 peer tgv 2>&1 | grep -E "mgpu_check|Error|error"
timeout 200 $TR tools/mgpu_check.py 128 3 peer rk4_f32 2>&1 | grep -E "mgpu_check|Error|error"
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1
timeout 600 $TR bench.py --gpus $N --grid 1024 --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1
} 2>&1 | tee gpurun_out/r2_mgpu8.log
