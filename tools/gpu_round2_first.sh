#!/bin/bash
# First GPU call of the next round: see the xfail-guarded tests on hardware, re-measure the bench line, run the
# measurements round 1 did not get to.  usage (under gpurun, one GPU): bash tools/gpu_round2_first.sh
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rxX > gpurun_out/r2_pytest.log 2>&1; tail -8 gpurun_out/r2_pytest.log
timeout 300 python tools/check_diagnostics.py 2>&1 | tail -4 | tee gpurun_out/r2_diagnostics.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; cut -c1-400 gpurun_out/r2_bench.json
timeout 300 python tools/perf_cases.py 20 2>&1 | grep perf_case | tee gpurun_out/r2_perf_cases.log
bash tools/gpu_fp64_probe.sh 2>&1 | tail -4
