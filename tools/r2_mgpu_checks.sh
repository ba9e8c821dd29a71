#!/bin/bash
# Multi-GPU checks on one box with N >= 2 GPUs (gpurun --gpus N -- 'bash tools/r2_mgpu_checks.sh [world]'):
# the native stepper (cpic_mgpu_*) against the oracle on `world` ranks over peer memory -- plain, and with one rank held
# back before every send (race detector for the landing buffers) -- then over NCCL, then bench lines on both transports.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
W=${1:-2}
{
nvidia-smi topo -m 2>&1 | head -$((W + 1)) | cut -c1-160
for SK in 0 500; do
  echo "== test_mgpu on $W ranks, peer memory required, rank 1 held back $SK us before every send"
  CPIC_TEST_WORLD=$W CPIC_REQUIRE_P2P=1 CPIC_P2P_TIMEOUT_S=20 CPIC_P2P_SKEW_US=$SK timeout 600 python -m pytest tests/test_mgpu.py -x -q -m gpu -s 2>&1 | grep -v "^$\|NCCL version" | tail -6
done
echo "== test_mgpu on $W ranks, NCCL forced"
CPIC_TEST_WORLD=$W CPIC_MGPU_P2P=0 timeout 600 python -m pytest tests/test_mgpu.py -x -q -m gpu -k "slab" 2>&1 | tail -2
for P in 1 0; do
  echo "== bench N=$W CPIC_MGPU_P2P=$P (256x256x$((32 * W)): 32-plane slabs)"
  CPIC_MGPU_P2P=$P timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 2961$P \
     bench.py --gpus $W --steps 20 --warmup 4 --no-e2e --grid 256 256 $((32 * W)) 2> gpurun_out/mgpu_p$P.err > gpurun_out/mgpu_p$P.json
  python - <<PY
import json
for l in open("gpurun_out/mgpu_p$P.json"):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]; p = d.get("parity") or {}
        print("N=%d  %.3f ms/step  push %.3f ms  non-push %.3f ms  launches %d  parity ok=%s E=%.6g B=%.6g migrated=%s  [%s]" % (
            d["n_gpus"], d["ms_per_step"], r["ms_per_launch"], d["ms_per_step"] - r["ms_per_launch"], d["gpu_launches"], p.get("ok"),
            p.get("e_energy", 0), p.get("b_energy", 0), p.get("migrated"), d["config"]["parallelism"][:110]))
PY
done
} 2>&1 | tee gpurun_out/mgpu_checks.log
