#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_chunks_verify.log; : > $L
timeout 300 python tools/perf_cases.py 20 2>&1 | grep perf_case >> $L
timeout 300 python tools/perf_cases.py 20 f32 2>&1 | grep perf_case >> $L
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 >> $L
cat $L
