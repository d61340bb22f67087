#!/bin/bash
mkdir -p gpurun_out
# memcheck of the stage kernels on small grids (ragged tiles, several z chunks), every scheme, both precisions, fourth and fifth generation
for a in 40,4,4,ls3 40,3,2,rk4 40,2,2,kutta 40,4,4,ls3,f32 40,1,1,rk4,f32 40,3,3,kutta,f32; do
  echo "== memcheck $a"; CUDNS_DUO=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/quick_perf.py $a 2>&1 | grep -E "ERROR SUMMARY|Invalid|error|n=" | head -5
done 2>&1 | tee gpurun_out/r2q_memcheck.log
echo "== memcheck fast kernel"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/quick_perf.py 40,4,4,ls3 2>&1 | grep -E "ERROR SUMMARY|Invalid|n=" | head -3 | tee -a gpurun_out/r2q_memcheck.log
timeout 600 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q -k bench_size 2>&1 | tail -3
# the reference's own GPU build through bench.py's in-job leg (5 vs 45 steps)
timeout 900 python -c "
import bench, json
print(json.dumps(bench.ref_gpu_baseline(512, 'ls3')))" | tee gpurun_out/r2q_ref_gpu_baseline.json
