#!/bin/bash
# tools/build_variant.sh <name> <extra nvcc flags...>: build a tuning variant of libcudns into gpurun-shipped
# build_var/<name>.so (select it with CUDNS_LIB=build_var/<name>.so)
set -e
NAME=$1; shift
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
W=/tmp/cudns_var/$NAME; D=$W/cudanavierstokes_b200/csrc; mkdir -p $D $W/include $ROOT/build_var
cp $ROOT/cudanavierstokes_b200/csrc/*.cu $ROOT/cudanavierstokes_b200/csrc/*.cpp $ROOT/cudanavierstokes_b200/csrc/*.h $ROOT/cudanavierstokes_b200/csrc/*.inc $ROOT/cudanavierstokes_b200/csrc/Makefile $D/
cp $ROOT/include/cudns.h $W/include/
make -s -j8 -C $D EXTRA="$*" > $D/make.log 2>&1 || { tail -20 $D/make.log; exit 1; }
cp $D/../libcudns.so $ROOT/build_var/$NAME.so
grep -A3 "Compiling.*stage_kernelILi4ELi4ELi[0-9]*ELi8ELb0" $D/stage_lean_s4.ptxas.log | grep -E "spill|Used"
