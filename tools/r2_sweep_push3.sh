#!/bin/bash
# k_push3 work-order knobs at the benched size: rows per z-block of the chunk order (CPIC_PUSH3_YBLOCK), cells per chunk
# (CPIC_PUSH3_CH), persistent grid size (CPIC_PUSH_GRID)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for nz in 256 64; do
  for yb in 2 4 8 16 32 64 300; do
    printf "256x256x%-3d yblock=%-3d  " $nz $yb
    CPIC_PUSH3_YBLOCK=$yb timeout 200 python tools/probe_reorder.py 256 256 $nz 64 6 reorder 2>&1 | tail -1
  done
  for ch in 86 129 258; do
    printf "256x256x%-3d ch=%-3d      " $nz $ch
    CPIC_PUSH3_CH=$ch timeout 200 python tools/probe_reorder.py 256 256 $nz 64 6 reorder 2>&1 | tail -1
  done
done
for gr in 296 444 592 888; do
  printf "256x256x64  grid=%-3d    " $gr
  CPIC_PUSH_GRID=$gr timeout 200 python tools/probe_reorder.py 256 256 64 64 6 reorder 2>&1 | tail -1
done
} 2>&1 | tee gpurun_out/c22.log
