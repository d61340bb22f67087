#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_diagnostics.py tests/test_post.py -m gpu -q > gpurun_out/r2k_pytest.log 2>&1; tail -3 gpurun_out/r2k_pytest.log
CUDNS_DUO=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/r2k_pytest_duo.log 2>&1; tail -3 gpurun_out/r2k_pytest_duo.log
timeout 600 python -m pytest tests/test_gpu_f32.py -m gpu -q -s > gpurun_out/r2k_pytest_f32.log 2>&1; grep "f32\|passed\|failed" gpurun_out/r2k_pytest_f32.log | tail -30
(echo "== fast (gen 4) ls3"; timeout 300 python tools/quick_perf.py 512,4,4 2>&1 | grep -v advance
echo "== duo"; CUDNS_DUO=1 timeout 300 python tools/quick_perf.py 512,4,4 512,4,4,rk4 512,4,4,kutta 2>&1 | grep -v advance
echo "== f32"; timeout 300 python tools/quick_perf.py 512,4,4,ls3,f32 512,4,4,rk4,f32 512,3,3,ls3,f32 2>&1) | tee gpurun_out/r2k_quick_perf.log
export CUDNS_DUO=1
bash tools/gpu_variants.sh "512,4,4 512,4,4,rk4" pf0 pf0s32
cp gpurun_out/variants.log gpurun_out/r2k_variants.log
ncu --set full --clock-control none --import-source on -k regex:stage_kernel -s 6 -c 1 -f -o gpurun_out/r2k_duo_f32_full python tools/quick_perf.py 512,4,4,ls3,f32 > gpurun_out/r2k_duo_f32_full.log 2>&1
