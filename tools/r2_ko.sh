#!/bin/bash
# Knock-out study of k_push3 (git apply tools/variants/push3_knockouts_dual_latepf.patch; make -C cabanapic_b200/csrc var NAME=ko
# DEFS=-DPUSH3_KO=1 -> libcabanapic_b200_ko.so; git apply -R afterwards): five normal steps, then ONE step with
# parts of the kernel disabled (wrong results, timing only).  Bits: 1 stayer record stores, 2 movers (list + drain),
# 4 foreigners' direct reductions, 8 segmented sum of the deposit rows, 16 gather always from the staged chunk,
# 32 slot claims, 64 the drain's reductions, 128 the drain's record store + histogram atomic.
cd "$(dirname "$0")/.."
for ko in 0 1 2 4 8 16 32 64 128 3 12 76 79 255; do
  printf "ko=%3d  " $ko
  CPIC_LIB=$PWD/cabanapic_b200/libcabanapic_b200_ko.so CPIC_KO_LAST=$ko timeout 120 python tools/probe_reorder.py 256 256 64 64 6 reorder 2>&1 | tail -1
done
