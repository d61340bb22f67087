#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:stage_kernel -s 6 -c 1 -f -o gpurun_out/r2_fast8_full python tools/quick_perf.py 512,4,4 > gpurun_out/r2_fast8_full.log 2>&1
tail -2 gpurun_out/r2_fast8_full.log
