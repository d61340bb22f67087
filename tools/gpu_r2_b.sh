#!/bin/bash
# round 2: the duo kernel on hardware -- parity first, then timing against the fourth generation
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -rxXs > gpurun_out/r2b_pytest.log 2>&1; tail -15 gpurun_out/r2b_pytest.log
timeout 300 python tools/quick_perf.py 512,4,4 512,4,4,rk4 512,4,4,kutta 256,4,4 512,3,3 512,2,2 2>&1 | tee gpurun_out/r2b_quick_perf.log
CUDNS_DUO=0 timeout 300 python tools/quick_perf.py 512,4,4 512,4,4,rk4 2>&1 | tee gpurun_out/r2b_quick_perf_gen4.log
