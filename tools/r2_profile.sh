#!/bin/bash
# final-shape captures: ncu --set full of k_push3 at 256x256x64 and at the benched 256^3, and the launch list of bench.py
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_push3 -s 4 -c 1 -f -o gpurun_out/profile_push3_256x256x64 \
  python tools/probe_reorder.py 256 256 64 64 6 reorder 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_push3 -s 4 -c 1 -f -o gpurun_out/profile_push3_256cube \
  python tools/probe_reorder.py 256 256 256 64 6 reorder 2>&1 | tail -1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/profile_launches_bench_c5.csv \
  python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/profile_bench_under_ncu.json 2> gpurun_out/profile_bench_under_ncu.err
tail -3 gpurun_out/profile_launches_bench_c5.csv | cut -c1-200
} 2>&1 | tee gpurun_out/profile.log
