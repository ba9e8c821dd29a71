#!/bin/bash
# native multi-GPU layer: parity tests on the GPUs of the box, then bench at N = number of GPUs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
{
echo "== test_mgpu on $N GPU(s)"
timeout 600 python -m pytest tests/test_mgpu.py -x -q -m gpu 2>&1 | tail -15
if [ "$N" -gt 1 ]; then
  echo "== bench N=$N native"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 4 2> gpurun_out/c9_bench.err | tee gpurun_out/c9_bench_n$N.json | cut -c1-1800
  tail -5 gpurun_out/c9_bench.err
  echo "== bench N=$N python stepper (round 1)"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 4 --python-stepper --no-e2e 2> gpurun_out/c9_bench_py.err | tee gpurun_out/c9_bench_py_n$N.json | cut -c1-600
fi
} 2>&1 | tee gpurun_out/c9_n$N.log
