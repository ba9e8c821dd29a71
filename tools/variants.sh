#!/bin/bash
# developer sweep: reordering-push / in-place push time for every libcabanapic_b200*.so variant in the package dir
# usage: tools/variants.sh nx ny nz [mode]
for lib in cabanapic_b200/libcabanapic_b200*.so; do
  echo "== $lib"
  CPIC_LIB=$PWD/$lib python tools/probe_reorder.py ${1:-256} ${2:-256} ${3:-64} 64 3 ${4:-both} 2>&1 | tail -4
done
