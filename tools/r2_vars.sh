#!/bin/bash
# time the steady-state reordering push of every variant library given (names after libcabanapic_b200_), then the product
cd "$(dirname "$0")/.."
for v in "$@" ""; do
  lib=cabanapic_b200/libcabanapic_b200${v:+_$v}.so
  [ -f "$lib" ] || continue
  echo "== ${v:-product}"
  CPIC_LIB=$PWD/$lib timeout 120 python tools/probe_reorder.py 256 256 64 64 6 reorder 2>&1 | tail -2
done
