#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r2_final3_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python tools/perf_cases.py 20 2>&1 | grep perf_case | tee gpurun_out/r2_final3_perf_cases.log
timeout 300 python tools/perf_cases.py 20 f32 2>&1 | grep perf_case | tee -a gpurun_out/r2_final3_perf_cases.log
