#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_f32.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r2_lean32_f32tests.log; cat gpurun_out/r2_lean32_f32tests.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
