#!/bin/bash
# 2-GPU session (run under gpurun --gpus 2): the multi-rank GPU tests, multi == single parity (both transports, both precisions), bench lines
N=2; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
{
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q 2>&1 | tail -3
timeout 300 $TR tools/mgpu_check.py 64 5 peer tgv 2>&1 | grep -E "mgpu_check|Error|error"
timeout 300 $TR tools/mgpu_check.py 64 5 nccl tgv 2>&1 | grep -E "mgpu_check|Error|error"
timeout 300 $TR tools/mgpu_check.py 64 4 peer kutta 2>&1 | grep -E "mgpu_check|Error|error"
timeout 300 $TR tools/mgpu_check.py 64 5 peer tgv_f32 2>&1 | grep -E "mgpu_check|Error|error"
timeout 300 $TR tools/mgpu_check.py 64 5 nccl tgv_f32 2>&1 | grep -E "mgpu_check|Error|error"
timeout 300 $TR tools/mgpu_check.py 64 4 peer rk4_f32 2>&1 | grep -E "mgpu_check|Error|error"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1
} 2>&1 | tee gpurun_out/r2_mgpu2.log
