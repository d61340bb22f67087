#!/bin/bash
# Run every oracle/_ref/ns_<cfg> on the GPU box and collect its outputs under
# gpurun_out/ref/<cfg>/ (fields/*.bin, solution.txt, Grid.txt, stdout).  CUDA_LAUNCH_BLOCKING=1
# serialises the reference's kernel launches, which removes its two latent inter-stream races
#  without touching its code, so the outputs are deterministic.
set -uo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OUT="$ROOT/gpurun_out/ref"
mkdir -p "$OUT"
for bin in "$HERE"/../_ref/ns_*; do
  name=$(basename "$bin"); name=${name#ns_}
  d="$OUT/$name"; rm -rf "$d"; mkdir -p "$d/fields" "$d/blasius1D"
  cp "$ROOT"/tests/golden/blasius1D/*.bin "$d/blasius1D/" 2>/dev/null || true
  ( cd "$d" && CUDA_LAUNCH_BLOCKING=1 timeout 300 "$bin" > stdout.txt 2>&1; echo "exit $?" >> stdout.txt )
  echo "== $name: $(tail -1 "$d/stdout.txt"), $(ls "$d/fields" | wc -l) files"
  rm -rf "$d/blasius1D"
done
