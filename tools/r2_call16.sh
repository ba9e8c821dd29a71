#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
CPIC_TEST_WORLD=4 CPIC_REQUIRE_P2P=1 CPIC_P2P_TIMEOUT_S=20 timeout 400 python -m pytest tests/test_mgpu.py -x -q -m gpu -k "graph_replay" 2>&1 | tail -60
echo "== NCCL forced"
CPIC_TEST_WORLD=4 CPIC_MGPU_P2P=0 timeout 400 python -m pytest tests/test_mgpu.py -x -q -m gpu -k "graph_replay" 2>&1 | tail -30
} 2>&1 | tee gpurun_out/c16.log
