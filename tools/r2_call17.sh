#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for SK in 0 500; do
echo "== test_mgpu on 4 ranks (grid 12x10x32), peer memory, rank 1 held back $SK us before every send"
CPIC_TEST_WORLD=4 CPIC_REQUIRE_P2P=1 CPIC_P2P_TIMEOUT_S=20 CPIC_P2P_SKEW_US=$SK timeout 500 python -m pytest tests/test_mgpu.py -x -q -m gpu -s 2>&1 | grep -v "^$\|NCCL version" | tail -8
done
} 2>&1 | tee gpurun_out/c17.log
