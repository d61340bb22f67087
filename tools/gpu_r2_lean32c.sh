#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_lean32_ctas.log; : > $L
for rep in 1 2; do
echo "== LEAN_CTAS=1 (rep $rep)" >> $L
CUDNS_LIB=build_var/lean_cta1.so timeout 300 python tools/perf_cases.py 20 f32 2>&1 | grep perf_case >> $L
echo "== LEAN_CTAS=2 (rep $rep)" >> $L
timeout 300 python tools/perf_cases.py 20 f32 2>&1 | grep perf_case >> $L
done
timeout 600 python -m pytest tests/test_gpu_f32.py tests/test_gpu_multirank.py -m gpu -q -x 2>&1 | tail -3 >> $L
cat $L
