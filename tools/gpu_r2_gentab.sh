#!/bin/bash
# A/B inside one job: stretched-grid tables and sponge references read from shared memory (staged per CTA / prefetched per plane)
# instead of global loads placed right in front of their use
mkdir -p gpurun_out
L=gpurun_out/r2_gen_smem_tables.log; : > $L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_f32.py tests/test_gpu_multirank.py tests/test_gpu_baseline_configs.py -m gpu -q -x 2>&1 | tail -3 >> $L
for rep in 1 2; do
echo "== global loads (rep $rep)" >> $L
CUDNS_LIB=build_var/gen_prev.so timeout 300 python tools/perf_cases.py 20 2>&1 | grep perf_case >> $L
CUDNS_LIB=build_var/gen_prev.so timeout 300 python tools/perf_cases.py 20 f32 2>&1 | grep perf_case >> $L
echo "== shared-memory tables (rep $rep)" >> $L
timeout 300 python tools/perf_cases.py 20 2>&1 | grep perf_case >> $L
timeout 300 python tools/perf_cases.py 20 f32 2>&1 | grep perf_case >> $L
done
cat $L
