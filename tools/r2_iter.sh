#!/bin/bash
# Round-2 kernel iteration on a GPU box: quick parity of the reordering push, steady-state timing, one ncu capture.
#   gpurun -- 'bash tools/r2_iter.sh TAG [ncu]'
cd "$(dirname "$0")/.."
TAG=${1:-it}
mkdir -p gpurun_out
{
echo "== parity (k_push3 forced on small grids)"
CPIC_PUSH2_PRIV=0 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "push_reorder or sorted_steps or fused_step or slab" 2>&1 | tail -3
echo "== timing"
timeout 300 python tools/probe_reorder.py 256 256 64 64 7 reorder | tail -3
timeout 300 python tools/probe_reorder.py 256 256 256 64 6 reorder | tail -2
if [ "$2" = "ncu" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_push3 -s 4 -c 1 -f -o gpurun_out/${TAG}_push3_256x256x64 \
    python tools/probe_reorder.py 256 256 64 64 6 reorder 2>&1 | tail -1
fi
} 2>&1 | tee gpurun_out/${TAG}.log
