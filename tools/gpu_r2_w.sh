#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2w_pytest_all.log 2>&1; tail -3 gpurun_out/r2w_pytest_all.log; grep -n "^E " gpurun_out/r2w_pytest_all.log | head -8
timeout 600 python tools/perf_cases.py 20 2>&1 | tee gpurun_out/r2w_perf_cases.log
(timeout 300 python tools/quick_perf.py 512,4,4 256,4,4 512,2,2 2>&1 | grep -v advance
CUDNS_DUO=1 timeout 300 python tools/quick_perf.py 512,4,4 512,4,4,rk4 256,4,4,rk4 2>&1 | grep -v advance
timeout 300 python tools/quick_perf.py 512,4,4,ls3,f32 256,4,4,ls3,f32 2>&1 | grep -v advance) | tee gpurun_out/r2w_quick_perf.log
