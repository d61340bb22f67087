#!/bin/bash
# A/B inside one job: who sends the next plane bundle of the fourth-generation stage kernel (thread 0 after its z sums / the last
# warp to empty the ring tile), and mbarrier try_wait with a suspend-time hint
mkdir -p gpurun_out
L=gpurun_out/r2_lastissue.log; : > $L
for rep in 1 2 3; do
for v in fl_base fl_last fl_hint fl_last_hint; do
  echo "== $v (rep $rep)" >> $L
  CUDNS_LIB=build_var/$v.so timeout 200 python tools/quick_perf.py 512,4,4 2>&1 | grep -v advance >> $L
done; done
echo "== parity with fl_last_hint" >> $L
CUDNS_LIB=build_var/fl_last_hint.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2 >> $L
cat $L
