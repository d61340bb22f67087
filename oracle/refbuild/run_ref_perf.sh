#!/bin/bash
# Time the reference's own GPU build (oracle/_ref/perf*) on this box; writes gpurun_out/ref_perf.txt.
# Lines: <name> <nsteps> <total wall seconds printed by the reference>
set -uo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
mkdir -p "$ROOT/gpurun_out"
OUT="$ROOT/gpurun_out/ref_perf.txt"; : > "$OUT"
for bin in "$HERE"/../_ref/perf*; do
  name=$(basename "$bin")
  d=$(mktemp -d /tmp/refperf.XXXXXX); mkdir -p "$d/fields"
  ( cd "$d" && timeout 900 "$bin" > stdout.txt 2>&1 )
  tot=$(grep "The total time is" "$d/stdout.txt" | awk '{print $5}')
  echo "$name ${name##*_n} ${tot:-FAILED}" | tee -a "$OUT"
  tail -3 "$d/stdout.txt" >> "$ROOT/gpurun_out/ref_perf_${name}.log"
  rm -rf "$d"
done
