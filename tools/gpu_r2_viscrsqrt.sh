#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_visc_rsqrt.log; : > $L
for rep in 1 2; do
echo "== sqrt (rep $rep)" >> $L
timeout 300 python tools/perf_cases.py 20 2>&1 | grep perf_case >> $L
echo "== rsqrt (rep $rep)" >> $L
CUDNS_LIB=build_var/visc_rsqrt.so timeout 300 python tools/perf_cases.py 20 2>&1 | grep perf_case >> $L
done
CUDNS_LIB=build_var/visc_rsqrt.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_f32.py -m gpu -q -x 2>&1 | tail -2 >> $L
cat $L
