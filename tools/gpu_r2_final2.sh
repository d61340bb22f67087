#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2_final2_pytest.log; cat gpurun_out/r2_final2_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_final2_bench.json 2> gpurun_out/r2_final2_bench.err; tail -3 gpurun_out/r2_final2_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r2_final2_bench.json').readline())
print(json.dumps({k:d[k] for k in ('value','ms_per_step','gpu_launches','cpu_baseline','ref_gpu_baseline','e2e','clocks')}, indent=None)[:1500])
print('roofline', {k:(round(v,4) if isinstance(v,float) else v) for k,v in d['roofline'].items() if k not in ('kernel','kernel_ms_how','kernel_ms_burst_how')})
for k,v in d['schemes'].items(): print(k, v['value'], v['ms_per_step'])
"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 | cut -c1-500
