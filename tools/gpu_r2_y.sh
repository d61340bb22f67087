#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python tools/perf_cases.py 10 2>&1 | cut -c1-150
timeout 300 python tools/quick_perf.py 512,4,4 512,4,4,ls3,f32 256,4,4 2>&1 | grep -v advance
CUDNS_DUO=1 timeout 300 python tools/quick_perf.py 512,4,4,rk4 2>&1 | grep -v advance
for n in 2 3 4 6 7 8 14 16; do echo "== theta chunks $n"; CUDNS_THETA_ZCHUNKS=$n timeout 300 python tools/quick_perf.py 512,4,4 512,4,4,ls3,f32 2>&1 | grep -v advance | cut -c1-60; done
} | tee gpurun_out/r2y_chunk_model.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2y_pytest_all.log 2>&1; tail -3 gpurun_out/r2y_pytest_all.log
