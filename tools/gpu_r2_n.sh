#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/issue_probe tools/issue_probe.cu && /tmp/issue_probe | tee gpurun_out/r2n_issue_probe.log
# sustained step loop: fourth- vs fifth-generation kernel for the low-storage stages, same job, 30 timed steps each
for d in 0 1; do
  if [ $d = 1 ]; then export CUDNS_DUO=1; else unset CUDNS_DUO; fi
  timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e --no-ref-gpu --no-schemes > gpurun_out/r2n_bench_duo$d.json 2>/dev/null
  python -c "
import json; d=json.loads(open('gpurun_out/r2n_bench_duo$d.json').readline()); r=d['roofline']
print('duo=$d value %.0f ms/step %.3f kernel in-loop %.3f burst %.3f theta %.3f clocks %s' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['kernel_ms_burst'], r['theta_ms'], d['clocks']))"
done 2>&1 | tee gpurun_out/r2n_sustained_gen4_vs_duo.log
unset CUDNS_DUO
# launch list of the bench command (kernel shares of the step; times under ncu are cold-cache and serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2n_launches_512.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-ref-gpu --no-schemes > gpurun_out/r2n_bench_under_ncu.log 2>&1
python tools/launch_shares.py gpurun_out/r2n_launches_512.csv | tee gpurun_out/r2n_launch_shares_512.txt
# DRAM traffic of the RK4 stage shape (duo MODE 3)
ncu --set full --clock-control none --import-source on -k regex:stage_kernel -s 8 -c 1 -f -o gpurun_out/r2n_duo_rk4_full python tools/quick_perf.py 512,4,4,rk4 > gpurun_out/r2n_duo_rk4_full.log 2>&1
# 1024^3 on ONE GPU in single precision (90 GB of device state; needs ~90 GB of host memory for the initial condition)
avail=$(awk '/MemAvailable/ {print int($2/1048576)}' /proc/meminfo); echo "host MemAvailable ${avail} GB"
if [ "$avail" -gt 200 ]; then timeout 900 python tools/quick_perf.py 1024,4,4,ls3,f32 2>&1 | tee gpurun_out/r2n_f32_1024.log; fi
