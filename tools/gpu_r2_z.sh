#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_zz_driver.py tests/test_post.py -m gpu -q > gpurun_out/r2z_pytest.log 2>&1; tail -4 gpurun_out/r2z_pytest.log; grep -n "^E " gpurun_out/r2z_pytest.log | head -12
