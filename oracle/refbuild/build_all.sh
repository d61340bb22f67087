#!/bin/bash
# Build every reference configuration of configs.sh into oracle/_ref/ (needs /root/reference).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
source "$HERE/configs.sh"
for name in "${!CFG[@]}"; do
  if [ ! -x "$HERE/../_ref/ns_$name" ] || [ "${FORCE:-0}" = 1 ]; then
    "$HERE/build_ref.sh" "$name" ${CFG[$name]}
  fi
done
for name in "${!PERF[@]}"; do
  if [ ! -x "$HERE/../_ref/$name" ] || [ "${FORCE:-0}" = 1 ]; then
    "$HERE/build_ref.sh" "$name" ${PERF[$name]}
    mv "$HERE/../_ref/ns_$name" "$HERE/../_ref/$name"
  fi
done
