#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multirank.py tests/test_gpu_baseline_configs.py -m gpu -q -x > gpurun_out/r2h_pytest.log 2>&1; tail -5 gpurun_out/r2h_pytest.log
timeout 300 python tools/quick_perf.py 512,4,4 512,2,2 256,4,4 2>&1 | tee gpurun_out/r2h_quick_perf.log
CUDNS_THETA_TMA=0 timeout 300 python tools/quick_perf.py 512,4,4 2>&1 | grep theta
ncu --set full --clock-control none --import-source on -k regex:theta_tma -s 6 -c 1 -f -o gpurun_out/r2h_theta_full python tools/quick_perf.py 512,4,4 > gpurun_out/r2h_theta_full.log 2>&1
