#!/bin/bash
mkdir -p gpurun_out
export CUDNS_DUO=1
(echo "== base"; timeout 300 python tools/quick_perf.py 512,4,4 512,4,4,rk4 2>&1 | grep -v advance) | tee gpurun_out/r2r_variants.log
bash tools/gpu_variants.sh "512,4,4 512,4,4,rk4" rcp early early_rcp
cat gpurun_out/variants.log >> gpurun_out/r2r_variants.log
CUDNS_LIB=$PWD/build_var/early_rcp.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2r_pytest_early_rcp.log
