#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py --steps 10 --warmup 3 --no-ref-gpu 2> gpurun_out/r2_final_bench.err | grep '^{' > gpurun_out/r2_final_bench.json; python -c "
import json
d=json.loads(open('gpurun_out/r2_final_bench.json').readline())
print(d['value'], d['ms_per_step'], d['gpu_launches'], d['e2e']['value'], d['cpu_baseline']['value'], d['clocks'])
print({k:(round(v,4) if isinstance(v,float) else v) for k,v in d['roofline'].items() if k in ('frac','frac_burst','kernel_ms','kernel_ms_burst','theta_ms','whole_step_frac')})
for k,v in d['schemes'].items(): print(k, v['value'], v['ms_per_step'])
"
