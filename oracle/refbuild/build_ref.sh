#!/bin/bash
# Build the reference's own GPU binary (`ns`) for one compile-time configuration, from the
# sources where they lie under /root/reference/src.  Nothing from the reference is copied into
# the repository: the patched working copy lives in a mktemp directory and only the linked
# binary lands in oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun).
#
# Mechanical patches:
#   1. drop the four dead CUDA-dynamic-parallelism functions (src/cuda_math.cu:54-141 and their
#      prototypes src/cuda_math.h:22-25) -- device-side cudaDeviceSynchronize no longer exists;
#   2. single-rank mpi.h stand-in (oracle/refbuild/mpi_stub/mpi.h);
#   3. globals.h: only the #define values named on the command line are changed;
#   4. for mx > 390 the x-metric tables do not fit __constant__: flip the `#if mx<=546` guards.
#
# usage: build_ref.sh <name> KEY=VALUE ...     e.g.  build_ref.sh tgv32 mx_tot=32 my_tot=32 ...
set -euo pipefail
REF=${CUDNS_REFERENCE_DIR:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../_ref"
NAME=$1; shift
[ -d "$REF/src" ] || { echo "reference sources not present at $REF (GPU box?) -- nothing to build"; exit 0; }
mkdir -p "$OUT"
W=$(mktemp -d /tmp/cudns_refbuild.XXXXXX)
trap 'rm -rf "$W"' EXIT
cp "$REF"/src/*.cu "$REF"/src/*.h "$REF"/src/*.cpp "$W"/
rm -f "$W/cuda_derivs.cu"                                  # dead legacy file, not in the main Makefile
sed -i '54,141d' "$W/cuda_math.cu"
sed -i '/__device__ void \(volumeIntegral\|reduceToOne\|reduceToMax\|reduceToMin\)/d' "$W/cuda_math.h"
BIGMX=0
for kv in "$@"; do
  k=${kv%%=*}; v=${kv#*=}
  case $k in
    stretch|TwallTop|TwallBot) sed -i -E "s|^(const [a-z]+ $k *= *)[^;]*;|\1$v;|" "$W/globals.h" ;;
    *) grep -qE "^#define $k[[:space:]]" "$W/globals.h" || { echo "unknown globals.h key $k"; exit 1; }
       sed -i -E "s|^#define $k[[:space:]].*|#define $k $v|" "$W/globals.h" ;;
  esac
  if [ "$k" = mx_tot ] && [ "$v" -gt 390 ]; then BIGMX=1; fi
done
if [ $BIGMX = 1 ]; then
  sed -i 's/#if mx<=546/#if 0/' "$W/cuda_globals.h" "$W/cuda_utils.cu"
  sed -i 's/__constant__ myprec d_dx, d_dy, d_dz, d_d2x, d_d2y, d_d2z, d_x\[mx\], d_xp\[mx\], d_dxv\[mx\];/__constant__ myprec d_dx, d_dy, d_dz, d_d2x, d_d2y, d_d2z, d_x[mx], d_xp[mx], d_dxv[mx];/' "$W/cuda_utils.cu"
fi
cd "$W"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
CXX=/usr/bin/g++
FLAGS="-rdc=true -gencode arch=compute_100,code=sm_100 -O3 --use_fast_math -ccbin $CXX -Xcompiler -mcmodel=medium -I$HERE/mpi_stub -I. -w"
for f in cuda_utils cuda_math cuda_main cuda_rhs calc_stress sponge; do
  $NVCC $FLAGS -c $f.cu -o $f.o &
done
for f in main comm init; do
  $CXX -O2 -std=c++11 -mcmodel=medium -fpermissive -w -I"$HERE/mpi_stub" -I. -c $f.cpp -o $f.o &
done
wait
$NVCC -gencode arch=compute_100,code=sm_100 -dlink cuda_utils.o cuda_math.o cuda_main.o cuda_rhs.o calc_stress.o sponge.o -o dlink.o -lcudadevrt
$CXX -mcmodel=medium -o "$OUT/ns_$NAME" main.o comm.o init.o cuda_utils.o cuda_math.o cuda_main.o cuda_rhs.o calc_stress.o sponge.o dlink.o \
    -Wl,--no-relax -L/usr/local/cuda/lib64 -Wl,-rpath,/usr/local/cuda/lib64 -lcudart -lcudadevrt -ldl -lrt -lpthread
echo "built $OUT/ns_$NAME"
