#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_f32.py tests/test_gpu_parity.py -m gpu -q -s > gpurun_out/r2m_pytest.log 2>&1; grep "f32\|passed\|failed" gpurun_out/r2m_pytest.log | tail -20
(echo "== f32 packed"; timeout 300 python tools/quick_perf.py 512,4,4,ls3,f32 512,4,4,rk4,f32 512,3,3,ls3,f32 512,2,2,ls3,f32 2>&1 | grep -v advance) | tee gpurun_out/r2m_quick_perf.log
ncu --set full --clock-control none --import-source on -k regex:stage_kernel -s 6 -c 1 -f -o gpurun_out/r2m_duo_f32p_full python tools/quick_perf.py 512,4,4,ls3,f32 > gpurun_out/r2m_duo_f32p_full.log 2>&1
