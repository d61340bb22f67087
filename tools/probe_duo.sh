#!/bin/bash
# compile only duo::stage_kernel<4,4,2> with extra -D flags and print registers / spills / opcode histogram (no GPU needed)
cd /root/repo/cudanavierstokes_b200/csrc
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I. "$@" -Xptxas -v -c /tmp/probe/probe.cu -o /tmp/probe/probe.o 2>&1 | grep -A2 "Compiling entry" | grep "spill\|registers"
/root/repo/tools/sass_stats.sh /tmp/probe/probe.o stage_kernelILi4ELi4ELi2E 10
echo "spill instrs: $(grep -c 'LDL\|STL' /tmp/sass_body.txt)  MOVs: $(grep -c 'IMAD.MOV.U32\|^MOV' /tmp/sass_body.txt)"
