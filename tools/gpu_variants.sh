#!/bin/bash
# time several build_var/*.so variants: bash tools/gpu_variants.sh "<quick_perf args>" name1 name2 ...
ARGS=$1; shift; mkdir -p gpurun_out
for v in "$@"; do
  echo "== $v"; CUDNS_LIB=$PWD/build_var/$v.so timeout 300 python tools/quick_perf.py $ARGS 2>&1 | grep -v advance
done | tee gpurun_out/variants.log
