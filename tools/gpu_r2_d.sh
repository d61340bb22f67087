#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multirank.py -m gpu -q -x > gpurun_out/r2e_pytest.log 2>&1; tail -3 gpurun_out/r2e_pytest.log
timeout 300 python tools/quick_perf.py 512,4,4 512,4,4,rk4 2>&1 | tee gpurun_out/r2e_quick_perf.log
bash tools/ncu_stage.sh 512 r2e_duo4 stage_kernel
