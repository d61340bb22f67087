#!/bin/bash
# A/B inside one job: ring insert of plane k+S before / behind the warp's arrival on the plane barrier (lean kernel)
mkdir -p gpurun_out
L=gpurun_out/r2_lean_ring_late.log; : > $L
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_f32.py tests/test_gpu_multirank.py tests/test_gpu_baseline_configs.py -m gpu -q -x 2>&1 | tail -2 >> $L
echo "== ring insert before the arrival" >> $L
CUDNS_LIB=build_var/ring_early.so timeout 300 python tools/perf_cases.py 20 2>&1 | grep perf_case >> $L
CUDNS_LIB=build_var/ring_early.so timeout 300 python tools/perf_cases.py 20 f32 2>&1 | grep perf_case >> $L
echo "== ring insert behind the arrival" >> $L
timeout 300 python tools/perf_cases.py 20 2>&1 | grep perf_case >> $L
timeout 300 python tools/perf_cases.py 20 f32 2>&1 | grep perf_case >> $L
cat $L
