#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:stage_kernel -s 6 -c 1 -f -o gpurun_out/r2_gen_chan2_final_full python tools/perf_one.py chan2 > gpurun_out/r2_gen_chan2_final_full.log 2>&1; tail -2 gpurun_out/r2_gen_chan2_final_full.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:stage_kernel -s 6 -c 1 -f -o gpurun_out/r2_gen_bl_final_full python tools/perf_one.py bl > gpurun_out/r2_gen_bl_final_full.log 2>&1; tail -2 gpurun_out/r2_gen_bl_final_full.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:theta_tma -s 6 -c 1 -f -o gpurun_out/r2_theta_gen_bl_full python tools/perf_one.py bl > gpurun_out/r2_theta_gen_bl_full.log 2>&1; tail -2 gpurun_out/r2_theta_gen_bl_full.log
