#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q > gpurun_out/r2u_pytest_multirank.log 2>&1; tail -4 gpurun_out/r2u_pytest_multirank.log; grep -n "^E " gpurun_out/r2u_pytest_multirank.log | head -8
