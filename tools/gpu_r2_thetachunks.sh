#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_theta_march_chunks.log; : > $L
for n in default 6 8 12 16 24 32 48 64 96 128; do
  echo "== CUDNS_THETA_ZCHUNKS=$n" >> $L
  if [ $n = default ]; then unset CUDNS_THETA_ZCHUNKS; else export CUDNS_THETA_ZCHUNKS=$n; fi
  timeout 200 python tools/perf_cases.py 10 2>&1 | grep perf_case | sed 's/Gpts.*| theta/| theta/' >> $L
done
cat $L
