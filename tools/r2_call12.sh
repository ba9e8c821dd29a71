#!/bin/bash
# 2-GPU check of the peer-memory slab exchange: parity tests on both transports, then bench lines on both
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-c12}
{
nvidia-smi topo -m 2>&1 | head -8
echo "== test_mgpu, peer memory required"
CPIC_REQUIRE_P2P=1 CPIC_P2P_TIMEOUT_S=20 timeout 400 python -m pytest tests/test_mgpu.py -x -q -m gpu -s 2>&1 | grep -v "^$" | tail -12
echo "== test_mgpu, NCCL forced"
CPIC_MGPU_P2P=0 timeout 400 python -m pytest tests/test_mgpu.py -x -q -m gpu -k "slab" 2>&1 | tail -3
for P in 1 0; do
  echo "== bench N=2 CPIC_MGPU_P2P=$P"
  CPIC_MGPU_P2P=$P CPIC_P2P_TIMEOUT_S=60 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2961$P \
     bench.py --gpus 2 --steps 20 --warmup 4 --no-e2e --grid 256 256 64 2> gpurun_out/${TAG}_p$P.err > gpurun_out/${TAG}_p$P.json
  python - <<PY
import json
for l in open("gpurun_out/${TAG}_p$P.json"):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]; p = d.get("parity") or {}
        print("N=%d  %.3f ms/step  push %.3f ms  non-push %.3f ms  launches %d  parity ok=%s migrated=%s  [%s]" % (
            d["n_gpus"], d["ms_per_step"], r["ms_per_launch"], d["ms_per_step"] - r["ms_per_launch"], d["gpu_launches"], p.get("ok"), p.get("migrated"), d["config"]["parallelism"]))
PY
  grep -v "^\*\|OMP_NUM" gpurun_out/${TAG}_p$P.err | tail -3 | cut -c1-300
done
} 2>&1 | tee gpurun_out/${TAG}.log
