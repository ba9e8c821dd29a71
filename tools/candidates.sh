#!/bin/bash
# Build the kernel candidates that are in the tree behind compile-time flags (all off in the product) and, on a GPU
# box, check each for parity and time the steady-state reordering push next to the product library.
#   tools/candidates.sh build        (here: nvcc cross-compiles)
#   gpurun --timeout 600 -- 'bash tools/candidates.sh run'
set -e
cd "$(dirname "$0")/.."
CANDS=("sg:-DPUSH2_SGATHER=1" "dagg:-DPUSH2_DRAINAGG=1" "sgdagg:-DPUSH2_SGATHER=1 -DPUSH2_DRAINAGG=1")
if [ "$1" = "build" ]; then
  for c in "${CANDS[@]}"; do make -C cabanapic_b200/csrc var NAME="${c%%:*}" DEFS="${c#*:}"; done
  exit 0
fi
for lib in cabanapic_b200/libcabanapic_b200.so cabanapic_b200/libcabanapic_b200_{sg,dagg,sgdagg}.so; do
  [ -f "$lib" ] || continue
  echo "== $lib"
  CPIC_LIB=$PWD/$lib timeout 120 python tools/probe_reorder.py 256 256 64 64 4 reorder 2>&1 | tail -2
  CPIC_LIB=$PWD/$lib timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "push_reorder_teacher or sorted_steps or push_strict" 2>&1 | tail -1
done
