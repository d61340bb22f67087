#!/bin/bash
# Build the reference's post-processing tool (postproc/post.cpp + comm.cpp + init.cpp: Favre / Reynolds means and fluctuations over
# saved fields) for one compile-time configuration, from the sources where they lie under /root/reference.  A host-only program: it
# builds AND runs in this container.  Nothing from the reference is copied into the repository: the working copy lives in a mktemp
# directory and only the linked binary lands in oracle/_ref/ (git-ignored).
#
# Mechanical patches: the tool includes "../src_multiGPU/{globals,comm,main}.h", a directory the repository does not ship -- the
# working copy provides it from src/; single-rank mpi.h stand-in (oracle/refbuild/mpi_stub); globals.h: only the #define values named
# on the command line are changed (same keys as build_ref.sh); the tool predates the rename of the macro `gamma` to `gam`
# (src/globals.h:46): its two uses (post.cpp:266, init.cpp:110) are renamed.
#
# usage: build_ref_post.sh <name> KEY=VALUE ...   ->  oracle/_ref/post_<name>   (run it from a directory next to "fields/")
set -euo pipefail
REF=${CUDNS_REFERENCE_DIR:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../_ref"
NAME=$1; shift
[ -d "$REF/postproc" ] || { echo "reference sources not present at $REF (GPU box?) -- nothing to build"; exit 0; }
mkdir -p "$OUT"
W=$(mktemp -d /tmp/cudns_refpost.XXXXXX)
trap 'rm -rf "$W"' EXIT
mkdir -p "$W/src_multiGPU" "$W/postproc"
cp "$REF"/src/*.h "$W/src_multiGPU/"
cp "$REF"/postproc/*.cpp "$W/postproc/"
G="$W/src_multiGPU/globals.h"
for kv in "$@"; do
  k=${kv%%=*}; v=${kv#*=}
  case $k in
    stretch|TwallTop|TwallBot) sed -i -E "s|^(const [a-z]+ $k *= *)[^;]*;|\1$v;|" "$G" ;;
    *) grep -qE "^#define $k[[:space:]]" "$G" || { echo "unknown globals.h key $k"; exit 1; }
       sed -i -E "s|^#define $k[[:space:]].*|#define $k $v|" "$G" ;;
  esac
done
cd "$W/postproc"
sed -i -E 's/\bgamma\b/gam/g' post.cpp init.cpp
CXX=/usr/bin/g++
for f in post comm init; do
  $CXX -O2 -std=c++11 -mcmodel=large -fpermissive -w -I"$HERE/mpi_stub" -c $f.cpp -o $f.o &
done
wait
$CXX -mcmodel=large -o "$OUT/post_$NAME" post.o comm.o init.o -lm
echo "built $OUT/post_$NAME"
