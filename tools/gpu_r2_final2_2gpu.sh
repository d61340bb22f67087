#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x -k "torchrun or team or driver" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_final2_bench_2gpu.json 2> gpurun_out/r2_final2_bench_2gpu.err; tail -2 gpurun_out/r2_final2_bench_2gpu.err; python -c "
import json
d=json.loads(open('gpurun_out/r2_final2_bench_2gpu.json').readline())
print(d['value'], d['ms_per_step'], d['n_gpus'], d.get('scaling'))
for k,v in d['schemes'].items(): print(k, v['value'], v['ms_per_step'])
"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/mgpu_check.py 128 5 peer tgv_s3v2_visc07 2>&1 | tail -4
