#!/bin/bash
# full GPU test suite + a short single-GPU bench line (with parity block and side lines)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "== bench N=1 (short)"
timeout 900 python bench.py --steps 6 --warmup 3 2> gpurun_out/c8_bench.err | tee gpurun_out/c8_bench_n1.json | cut -c1-3000
tail -5 gpurun_out/c8_bench.err
} 2>&1 | tee gpurun_out/c8.log
