#!/bin/bash
# run under gpurun: FP64 pipe vs FP64 tensor-core throughput AND power (see tools/fp64_probe.cu)
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_probe tools/fp64_probe.cu || exit 1
nvidia-smi --query-gpu=power.draw,clocks.sm --format=csv,noheader -lms 200 > gpurun_out/fp64_probe_power.csv &
SMI=$!
/tmp/fp64_probe 4 | tee gpurun_out/fp64_probe.log
kill $SMI
python - <<'PY'
rows = [l.strip().split(",") for l in open("gpurun_out/fp64_probe_power.csv") if "W" in l]
w = [float(r[0].split()[0]) for r in rows]; n = len(w) // 2
print("power: DFMA leg max %.0f W, DMMA leg max %.0f W (samples %d)" % (max(w[:n]) if n else 0, max(w[n:]) if n else 0, len(w)))
PY
