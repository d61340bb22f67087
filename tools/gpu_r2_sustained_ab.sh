#!/bin/bash
# sustained (power-capped) step loop: fourth generation (default) vs fifth generation (CUDNS_DUO=1) for low-storage RK3, same job
mkdir -p gpurun_out
L=gpurun_out/r2_sustained_fast_vs_duo.log; : > $L
for rep in 1 2; do
for v in default duo; do
  if [ $v = duo ]; then export CUDNS_DUO=1; else unset CUDNS_DUO; fi
  timeout 300 python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-schemes --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']
        print('$v rep $rep: value %.0f ms/step %.3f kernel_ms %.3f burst %.3f theta %.3f clocks %s W %s' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['kernel_ms_burst'], r['theta_ms'], d['clocks']['sm_mhz'], d['clocks'].get('power_w_max')))
" >> $L
done; done
cat $L
