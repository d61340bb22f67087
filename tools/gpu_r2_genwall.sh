#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_gen_wall_per_warp.log; : > $L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_f32.py tests/test_gpu_multirank.py tests/test_gpu_baseline_configs.py -m gpu -q -x 2>&1 | tail -3 >> $L
for rep in 1 2; do
timeout 300 python tools/perf_cases.py 20 2>&1 | grep perf_case >> $L
timeout 300 python tools/perf_cases.py 20 f32 2>&1 | grep perf_case >> $L
done
cat $L
