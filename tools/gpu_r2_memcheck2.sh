#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_memcheck_lean.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_f32.py tests/test_gpu_parity.py -m gpu -q -x -k "walls or lean_kernel or reference_gpu_binary or channel" > $L 2>&1
echo "exit $?" >> $L
grep -E "ERROR SUMMARY|passed|failed|exit|Invalid|error" $L | head -20
