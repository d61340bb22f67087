#!/bin/bash
mkdir -p gpurun_out
for D in 1 0; do
  echo "CUDNS_DUO=$D"
  CUDNS_DUO=$D timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/r2g_bench_duo$D.err | tee gpurun_out/r2g_bench_duo$D.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('value %.0f ms/step %.2f kernel_ms %.3f theta %.3f clocks %s power %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['theta_ms'], d['clocks']['sm_mhz'], d['clocks'].get('power_w_max')))"
done
CUDNS_DUO=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --scheme rk4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('rk4 value %.0f ms/step %.2f kernel_ms %.3f clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['clocks']['sm_mhz']))"
