#!/bin/bash
# usage: tools/ncu_stage.sh <n> <tag>   -- ncu --set full capture of the stage kernel + launch list (run under gpurun)
N=${1:-256}; TAG=${2:-stage}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:stage_kernel -s 6 -c 1 -f -o gpurun_out/${TAG}_full \
    python tools/quick_perf.py $N,4,4 > gpurun_out/${TAG}_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --n $N --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_launches.log 2>&1
