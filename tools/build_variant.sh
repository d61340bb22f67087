#!/bin/bash
# tools/build_variant.sh <name> <extra nvcc flags...>: build a tuning variant of libcudns into gpurun-shipped build_var/<name>.so
set -e
NAME=$1; shift
cd "$(dirname "$0")/../cudanavierstokes_b200/csrc"
D=../../build_var/$NAME; mkdir -p $D
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC -Xptxas -v $@"
$NV -c kernels.cu -o $D/kernels.o 2>/dev/null &
$NV -c stage_tmem.cu -o $D/stage_tmem.o 2> $D/stage_tmem.ptxas.log &
$NV -c api.cu -o $D/api.o 2>/dev/null &
$NV -x cu -c host_setup.cpp -o $D/host_setup.o 2>/dev/null &
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o ../../build_var/$NAME.so $D/kernels.o $D/stage_tmem.o $D/api.o $D/host_setup.o -lcudart_static -ldl -lrt -lpthread
grep -A3 "stage_kernelILi4ELi4ELi[0-9]*ELi8" $D/stage_tmem.ptxas.log | grep -E "spill|Used"
