#!/bin/bash
cd "$(dirname "$0")/.."
{
echo "== parity (k_push3 forced on small grids), new block shape"
CPIC_PUSH2_PRIV=0 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "push_reorder or sorted_steps or fused_step or slab or reflect or species or odd_count" 2>&1 | tail -2
bash tools/r2_vars.sh dual latepf
timeout 300 python tools/probe_reorder.py 256 256 256 64 6 reorder | tail -1
} 2>&1 | tee gpurun_out/c26.log
