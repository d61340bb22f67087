#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_theta_tma_gen.log; : > $L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_f32.py tests/test_gpu_multirank.py tests/test_gpu_baseline_configs.py tests/test_zz_diagnostics.py tests/test_zz_driver.py tests/test_post.py -m gpu -q -x 2>&1 | tail -3 >> $L
for rep in 1 2; do
echo "== cp.async dilatation pass (CUDNS_THETA_TMA=0), rep $rep" >> $L
CUDNS_THETA_TMA=0 timeout 300 python tools/perf_cases.py 20 2>&1 | grep perf_case >> $L
echo "== TMA dilatation pass, rep $rep" >> $L
timeout 300 python tools/perf_cases.py 20 2>&1 | grep perf_case >> $L
done
echo "== single precision: cp.async / TMA" >> $L
CUDNS_THETA_TMA=0 timeout 300 python tools/perf_cases.py 20 f32 2>&1 | grep perf_case >> $L
timeout 300 python tools/perf_cases.py 20 f32 2>&1 | grep perf_case >> $L
cat $L
