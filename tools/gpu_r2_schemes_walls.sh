#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "three_register_schemes" 2>&1 | grep -v "^$" | tail -14 | tee gpurun_out/r2_schemes_walls.log
