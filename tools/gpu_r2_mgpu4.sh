#!/bin/bash
N=4; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
{
timeout 200 $TR tools/mgpu_check.py 128 3 peer tgv 2>&1 | grep -E "mgpu_check|Error|error"
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1
timeout 300 cudanavierstokes_b200/cudns_run case=tgv mx=512 my=512 mz=512 stencilSize=4 nsteps=100 nfiles=1 ngpus=4 async_io=1 outdir=/tmp/run4 xdmf=0 2>&1 | grep -E "cudns_run|total time|file number"
} 2>&1 | tee gpurun_out/r2_mgpu4.log
