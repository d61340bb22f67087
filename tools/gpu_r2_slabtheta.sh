#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_slab_theta_chunks.log; : > $L
for mz in 64 128; do for n in default 1 2 3 4; do
  if [ $n = default ]; then unset CUDNS_THETA_ZCHUNKS; else export CUDNS_THETA_ZCHUNKS=$n; fi
  timeout 120 python tools/slab_perf.py 512 512 $mz 2>&1 | grep slab >> $L
done; done
unset CUDNS_THETA_ZCHUNKS
for n in default 1 2 3; do
  if [ $n = default ]; then unset CUDNS_ZCHUNKS; else export CUDNS_ZCHUNKS=$n; fi
  timeout 120 python tools/slab_perf.py 512 512 64 2>&1 | grep slab >> $L
done
cat $L
