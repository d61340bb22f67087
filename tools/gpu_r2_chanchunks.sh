#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_channel_stage_chunks.log; : > $L
for n in default 5 7 9 11 12 13 16 24; do
  echo "== CUDNS_ZCHUNKS=$n" >> $L
  if [ $n = default ]; then unset CUDNS_ZCHUNKS; else export CUDNS_ZCHUNKS=$n; fi
  timeout 200 python tools/perf_cases.py 10 f64 channel 2>&1 | grep perf_case | sed 's/Gpts.*| theta/| theta/' >> $L
done
unset CUDNS_ZCHUNKS
for n in default 25 37 50 62 74 100; do
  echo "== boundary layer CUDNS_ZCHUNKS=$n" >> $L
  if [ $n = default ]; then unset CUDNS_ZCHUNKS; else export CUDNS_ZCHUNKS=$n; fi
  timeout 200 python tools/perf_cases.py 10 f64 boundary 2>&1 | grep perf_case | sed 's/Gpts.*| theta/| theta/' >> $L
done
cat $L
