#!/bin/bash
# One GPU-box session: parity tests, quick per-kernel timing, bench line, ncu launch list + one full capture.
# usage (under gpurun): bash tools/gpu_session.sh <tag> [skip-tests]
TAG=${1:-s}; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
if [ "${2:-}" != "skip-tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
  tail -3 gpurun_out/${TAG}_pytest.log
fi
timeout 300 python tools/quick_perf.py 256,4,4 512,4,4 256,3,3 > gpurun_out/${TAG}_quick.log 2>&1; cat gpurun_out/${TAG}_quick.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv \
   python bench.py --n 256 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage_kernel -s 6 -c 1 -f -o gpurun_out/${TAG}_stage_full \
   python tools/quick_perf.py 512,4,4 > gpurun_out/${TAG}_stage_full.log 2>&1
ls -la gpurun_out | tail -20
