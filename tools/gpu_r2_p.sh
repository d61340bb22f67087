#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2p_pytest_all.log 2>&1; tail -3 gpurun_out/r2p_pytest_all.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; tail -3 gpurun_out/r2p_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r2p_bench.json').readline())
print(json.dumps({k:d[k] for k in ('value','ms_per_step','gpu_launches','cpu_baseline','ref_gpu_baseline','e2e','clocks')}, indent=None)[:1500])
print('roofline', {k:(round(v,4) if isinstance(v,float) else v) for k,v in d['roofline'].items() if k not in ('kernel','kernel_ms_how','kernel_ms_burst_how')})
for nm, s in d['schemes'].items():
    print(nm, round(s['value']), round(s['ms_per_step'],3), {k:(round(v,4) if isinstance(v,float) else v) for k,v in s['roofline'].items() if k in ('frac','frac_burst','kernel_ms','kernel_ms_burst','whole_step_frac','theta_ms','alg_bytes_per_point','traffic')})
"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
