#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q -k bench_size > gpurun_out/r2s_pytest_props.log 2>&1; tail -3 gpurun_out/r2s_pytest_props.log; grep -n "^E " gpurun_out/r2s_pytest_props.log | head -5
export CUDNS_DUO=1
(echo "== base (fast reciprocal)"; timeout 300 python tools/quick_perf.py 512,4,4 512,4,4,rk4 2>&1 | grep -v advance) | tee gpurun_out/r2s_variants.log
bash tools/gpu_variants.sh "512,4,4 512,4,4,rk4" rcp0
cat gpurun_out/variants.log >> gpurun_out/r2s_variants.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/r2s_pytest_duo.log 2>&1; tail -3 gpurun_out/r2s_pytest_duo.log
