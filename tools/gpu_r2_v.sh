#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/perf_cases.py 20 2>&1 | tee gpurun_out/r2v_perf_cases.log
# ncu of the general (walls / stretched grid / non-linear viscosity) stage kernel: channel 160x192x192 at 4th order
ncu --set full --clock-control none --import-source on -k regex:stage_kernel -s 10 -c 1 -f -o gpurun_out/r2v_lean_gen_channel_full python tools/perf_cases.py 3 > gpurun_out/r2v_lean_gen_channel_full.log 2>&1
