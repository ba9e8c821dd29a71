#!/bin/bash
# Round-2 GPU call 1: (a) size sweep of the steady-state reordering push with clocks sampled, (b) the two flagged
# candidates of round 1, (c) ncu --set full of k_push2 at the BENCHED size (256^3 x 64) and at 256x256x64.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
Q=index,timestamp,clocks.sm,clocks.max.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown
nvidia-smi --query-gpu=$Q --format=csv -lms 100 > gpurun_out/c1_clocks.csv &
SMI=$!
for nz in 32 64 128 256; do
  echo "== 256x256x$nz $(date +%T.%N)"
  timeout 300 python tools/probe_reorder.py 256 256 $nz 64 10 reorder
done > gpurun_out/c1_sizes.log 2>&1
kill $SMI
bash tools/candidates.sh run > gpurun_out/c1_cands.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_push2 -s 4 -c 1 -f -o gpurun_out/c1_push2_256cube \
  python tools/probe_reorder.py 256 256 256 64 6 reorder > gpurun_out/c1_ncu256.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_push2 -s 4 -c 1 -f -o gpurun_out/c1_push2_256x256x64 \
  python tools/probe_reorder.py 256 256 64 64 6 reorder > gpurun_out/c1_ncu64.log 2>&1
ls -la gpurun_out
tail -3 gpurun_out/c1_ncu256.log gpurun_out/c1_ncu64.log
cat gpurun_out/c1_sizes.log | grep -E "==|step [5-9]"
cat gpurun_out/c1_cands.log
