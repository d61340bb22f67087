#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_f32.py tests/test_gpu_multirank.py tests/test_gpu_baseline_configs.py tests/test_zz_diagnostics.py tests/test_zz_driver.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/perf_cases.py 20 2>&1 | grep perf_case | tee gpurun_out/r2_perf_cases_final.log
timeout 300 python tools/perf_cases.py 20 f32 2>&1 | grep perf_case | tee -a gpurun_out/r2_perf_cases_final.log
