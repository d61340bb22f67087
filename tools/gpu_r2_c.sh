#!/bin/bash
# duo kernel with the cp.async operand stash: parity of the duo tests, timing, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multirank.py tests/test_gpu_baseline_configs.py -m gpu -q -x > gpurun_out/r2c_pytest.log 2>&1; tail -5 gpurun_out/r2c_pytest.log
timeout 300 python tools/quick_perf.py 512,4,4 512,4,4,rk4 512,4,4,kutta 512,2,2 2>&1 | tee gpurun_out/r2c_quick_perf.log
bash tools/ncu_stage.sh 512 r2c_duo2 stage_kernel
