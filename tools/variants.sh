#!/bin/bash
# developer sweep: push time for every libcabanapic_b200_*.so variant lying in the package dir
for lib in cabanapic_b200/libcabanapic_b200*.so; do
  echo "== $lib"
  CPIC_LIB=$PWD/$lib python tools/probe.py ${1:-128} ${2:-128} ${3:-128} 64 2>&1 | grep -E "strict   warp|contract warp|step [014] since"
done
