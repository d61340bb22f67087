#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_gen_suspend_hint.log; : > $L
for rep in 1 2; do
for v in base gen_hint500 gen_hint2000; do
echo "== $v (rep $rep)" >> $L
if [ $v = base ]; then unset CUDNS_LIB; else export CUDNS_LIB=build_var/$v.so; fi
timeout 300 python tools/perf_cases.py 20 2>&1 | grep perf_case >> $L
done; done
unset CUDNS_LIB
cat $L
