#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:stage_kernel -s 6 -c 1 -f -o gpurun_out/r2_gen_bl_ty12_full python tools/perf_one.py bl > gpurun_out/r2_gen_bl_ty12_full.log 2>&1; tail -2 gpurun_out/r2_gen_bl_ty12_full.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:stage_kernel -s 6 -c 1 -f -o gpurun_out/r2_gen_bl_f32_full python tools/perf_one.py bl f32 > gpurun_out/r2_gen_bl_f32_full.log 2>&1; tail -2 gpurun_out/r2_gen_bl_f32_full.log
