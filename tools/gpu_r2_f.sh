#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2f_pytest.log 2>&1; tail -3 gpurun_out/r2f_pytest.log
for K in 0 1500 3000 6000; do echo "knob $K"; CUDNS_DUO_KNOB=$K timeout 300 python tools/quick_perf.py 512,4,4 2>&1 | grep rhs_stage; done | tee gpurun_out/r2f_knob.log
