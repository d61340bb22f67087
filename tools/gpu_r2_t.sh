#!/bin/bash
mkdir -p gpurun_out
export CUDNS_DUO=1
for rep in 1 2; do bash tools/gpu_variants.sh "512,4,4 512,4,4,rk4" pad nopad; cat gpurun_out/variants.log >> gpurun_out/r2t_variants.log; done
unset CUDNS_DUO
for rep in 1 2; do bash tools/gpu_variants.sh "512,4,4" pad nopad; cat gpurun_out/variants.log >> gpurun_out/r2t_variants.log; done
