#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2_lean32_fullsuite.log; cat gpurun_out/r2_lean32_fullsuite.log
timeout 300 python tools/perf_cases.py 20 2>&1 | grep perf_case > gpurun_out/r2_perf_cases_f64_f32.log
timeout 300 python tools/perf_cases.py 20 f32 2>&1 | grep perf_case >> gpurun_out/r2_perf_cases_f64_f32.log
cat gpurun_out/r2_perf_cases_f64_f32.log
