#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/issue_probe tools/issue_probe.cu && /tmp/issue_probe | tee gpurun_out/r2o_issue_probe.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_f32.py -m gpu -q > gpurun_out/r2o_pytest.log 2>&1; tail -3 gpurun_out/r2o_pytest.log
CUDNS_DUO=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/r2o_pytest_duo.log 2>&1; tail -3 gpurun_out/r2o_pytest_duo.log
(CUDNS_DUO=1 timeout 300 python tools/quick_perf.py 512,4,4 512,4,4,rk4 512,4,4,kutta 2>&1 | grep -v advance
timeout 300 python tools/quick_perf.py 512,4,4,ls3,f32 512,4,4,rk4,f32 512,4,4,kutta,f32 2>&1 | grep -v advance) | tee gpurun_out/r2o_quick_perf.log
python tools/check_diagnostics.py 2>&1 | tee gpurun_out/r2o_check_diagnostics.log
ncu --set full --clock-control none --import-source on -k regex:stage_kernel -s 8 -c 1 -f -o gpurun_out/r2o_duo_rk4_full python tools/quick_perf.py 512,4,4,rk4 > gpurun_out/r2o_duo_rk4_full.log 2>&1
