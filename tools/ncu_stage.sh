#!/bin/bash
# usage: tools/ncu_stage.sh <n> <tag> [kernel regex] [lib]  -- ncu --set full capture of one kernel (run under gpurun)
N=${1:-256}; TAG=${2:-stage}; K=${3:-stage_kernel}; LIB=${4:-}
mkdir -p gpurun_out
[ -n "$LIB" ] && export CUDNS_LIB=$PWD/$LIB
ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 1 -f -o gpurun_out/${TAG}_full \
    python tools/quick_perf.py $N,4,4 > gpurun_out/${TAG}_full.log 2>&1
