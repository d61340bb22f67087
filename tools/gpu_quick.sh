#!/bin/bash
# short GPU session: parity tests + per-kernel timing.  usage: bash tools/gpu_quick.sh <tag> [quick_perf args...]
TAG=${1:-q}; shift; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/quick_perf.py ${@:-256,4,4 512,4,4} > gpurun_out/${TAG}_quick.log 2>&1; cat gpurun_out/${TAG}_quick.log
