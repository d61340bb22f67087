#!/bin/bash
# A/B inside one job: ring tile on a barrier of its own (loaded first), halo'd plane waited for after the z sums
mkdir -p gpurun_out
L=gpurun_out/r2_splitbar.log; : > $L
export QP_REPS=20
for rep in 1 2 3; do
for v in fs_base fs_split fs_split_hint; do
  echo "== $v (rep $rep)" >> $L
  CUDNS_LIB=build_var/$v.so timeout 200 python tools/quick_perf.py 512,4,4 2>&1 | grep -v advance >> $L
done; done
echo "== parity with fs_split" >> $L
CUDNS_LIB=build_var/fs_split.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2 >> $L
cat $L
