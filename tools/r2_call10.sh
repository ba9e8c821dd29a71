#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_mgpu.py tests/test_facade.py -x -q -m gpu 2>&1 | tail -15
} 2>&1 | tee gpurun_out/c10.log
