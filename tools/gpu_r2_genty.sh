#!/bin/bash
# A/B inside one job: 9-quantity lean kernel (non-linear viscosity) with 8 vs 12 tile rows (= warps per SM) for s <= 3, FP64
mkdir -p gpurun_out
L=gpurun_out/r2_gen_ty12.log; : > $L
for rep in 1 2; do
echo "== TY 8 (rep $rep)" >> $L
CUDNS_LIB=build_var/gen_ty8.so timeout 300 python tools/perf_cases.py 20 2>&1 | grep perf_case >> $L
echo "== TY 12 (rep $rep)" >> $L
timeout 300 python tools/perf_cases.py 20 2>&1 | grep perf_case >> $L
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multirank.py tests/test_gpu_baseline_configs.py -m gpu -q -x 2>&1 | tail -3 >> $L
cat $L
