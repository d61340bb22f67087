#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2i_pytest.log 2>&1; tail -3 gpurun_out/r2i_pytest.log
timeout 300 python tools/quick_perf.py 512,4,4,rk4 512,4,4,kutta 2>&1 | tee gpurun_out/r2i_quick_perf.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; tail -3 gpurun_out/r2i_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r2i_bench.json').readline())
print(json.dumps({k:d[k] for k in ('value','ms_per_step','gpu_launches','cpu_baseline','ref_gpu_baseline','e2e','clocks')}, indent=None)[:1500])
print('roofline', {k:(round(v,4) if isinstance(v,float) else v) for k,v in d['roofline'].items() if k not in ('kernel','kernel_ms_how','kernel_ms_burst_how')})
print('rk4', d['schemes']['rk4']['value'], d['schemes']['rk4']['ms_per_step'], {k:(round(v,4) if isinstance(v,float) else v) for k,v in d['schemes']['rk4']['roofline'].items() if k in ('frac','frac_burst','kernel_ms','kernel_ms_burst','whole_step_frac','theta_ms')})
"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 | cut -c1-600
