#!/bin/bash
# A/B inside one job: x direction before / after the wait for the other warps' shares of the shared plane (lean kernel)
mkdir -p gpurun_out
L=gpurun_out/r2_lean_x_early.log; : > $L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_f32.py tests/test_gpu_multirank.py tests/test_gpu_baseline_configs.py -m gpu -q -x 2>&1 | tail -2 >> $L
for rep in 1 2; do
echo "== x after the wait (rep $rep)" >> $L
CUDNS_LIB=build_var/x_late.so timeout 300 python tools/perf_cases.py 20 2>&1 | grep perf_case >> $L
echo "== x before the wait (rep $rep)" >> $L
timeout 300 python tools/perf_cases.py 20 2>&1 | grep perf_case >> $L
done
echo "== single precision: after / before" >> $L
CUDNS_LIB=build_var/x_late.so timeout 300 python tools/perf_cases.py 20 f32 2>&1 | grep perf_case >> $L
timeout 300 python tools/perf_cases.py 20 f32 2>&1 | grep perf_case >> $L
cat $L
