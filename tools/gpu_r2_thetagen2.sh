#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_theta_tma_gen_chunks.log; : > $L
timeout 300 python tools/perf_cases.py 20 2>&1 | grep perf_case >> $L
timeout 300 python tools/perf_cases.py 20 f32 2>&1 | grep perf_case >> $L
for n in 4 8 16 32; do echo "== CUDNS_THETA_ZCHUNKS=$n" >> $L; CUDNS_THETA_ZCHUNKS=$n timeout 300 python tools/perf_cases.py 10 2>&1 | grep "perf_case channel" | sed 's/Gpts.*| theta/| theta/' >> $L; done
timeout 200 python tools/quick_perf.py 512,4,4 256,4,4 2>&1 | grep -v advance >> $L
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_f32.py tests/test_gpu_multirank.py -m gpu -q -x 2>&1 | tail -2 >> $L
cat $L
