#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== parity (k_push3 forced on small grids)"
CPIC_PUSH2_PRIV=0 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "push_reorder or sorted_steps or fused_step or slab or reflect or species" 2>&1 | tail -3
echo "== timing product"
timeout 300 python tools/probe_reorder.py 256 256 64 64 7 reorder | tail -2
timeout 300 python tools/probe_reorder.py 256 256 256 64 6 reorder | tail -2
echo "== knock-outs"
bash tools/r2_ko.sh
echo "== ncu 256^3"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_push3 -s 4 -c 1 -f -o gpurun_out/c11_push3_256cube \
  python tools/probe_reorder.py 256 256 256 64 6 reorder 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_push3 -s 4 -c 1 -f -o gpurun_out/c11_push3_256x256x64 \
  python tools/probe_reorder.py 256 256 64 64 6 reorder 2>&1 | tail -1
} 2>&1 | tee gpurun_out/c11.log
