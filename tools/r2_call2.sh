#!/bin/bash
# Round-2 GPU call 2: first run of k_push3 (block-owned cell ranges, TMA-staged interpolators): parity, then timing.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
K="push_reorder or sorted_steps or fused_step or slab or smoke or full_size_c5 or block_private or golden"
echo "== parity, k_push3 forced on small grids (CPIC_PUSH2_PRIV=0)"
CPIC_PUSH2_PRIV=0 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K" 2>&1 | tail -15
echo "== parity, default dispatch"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
echo "== timing k_push3"
for nz in 64 256; do timeout 300 python tools/probe_reorder.py 256 256 $nz 64 8 reorder; done
echo "== timing k_push2 (CPIC_PUSH3=0)"
CPIC_PUSH3=0 timeout 300 python tools/probe_reorder.py 256 256 64 64 6 reorder | tail -2
