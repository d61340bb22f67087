#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_gen_ring_table.log; : > $L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_f32.py tests/test_gpu_multirank.py tests/test_gpu_baseline_configs.py -m gpu -q -x 2>&1 | tail -3 >> $L
for rep in 1 2; do
echo "== ring slot table (rep $rep)" >> $L
timeout 300 python tools/perf_cases.py 20 2>&1 | grep perf_case >> $L
echo "== ring slot table + s = 3 stencil loop fully unrolled (rep $rep)" >> $L
CUDNS_LIB=build_var/gen_unroll3.so timeout 300 python tools/perf_cases.py 20 2>&1 | grep perf_case >> $L
done
timeout 300 python tools/perf_cases.py 20 f32 2>&1 | grep perf_case >> $L
cat $L
