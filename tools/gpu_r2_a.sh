#!/bin/bash
# round 2, first GPU call: parity holes (new tests, un-xfailed diagnostics), diagnostics log, full-size channel / boundary-layer throughput,
# fast kernel after the s0-spill fix, FP64 pipe / DMMA probe.   usage (under gpurun, one GPU): bash tools/gpu_r2_a.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > gpurun_out/r2a_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -rxXs --durations=15 > gpurun_out/r2a_pytest.log 2>&1; tail -30 gpurun_out/r2a_pytest.log
timeout 300 python tools/check_diagnostics.py > gpurun_out/r2a_diagnostics.log 2>&1; cat gpurun_out/r2a_diagnostics.log
timeout 300 python tools/quick_perf.py 512,4,4 256,4,4 2>&1 | tee gpurun_out/r2a_quick_perf.log
timeout 400 python tools/perf_cases.py 20 2>&1 | grep perf_case | tee gpurun_out/r2a_perf_cases.log
bash tools/gpu_fp64_probe.sh 2>&1 | tail -6 | tee gpurun_out/r2a_fp64_probe.log
