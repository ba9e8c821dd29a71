#!/bin/bash
# 1 / 2 / 4 / 8 GPU bench lines on ONE box (native C++ stepper), as the driver's scaling run does
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-scale}
NG=$(nvidia-smi -L | wc -l)
for N in 1 2 4 8; do
  [ "$N" -le "$NG" ] || continue
  echo "== N=$N"
  if [ "$N" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-extras --no-e2e 2> gpurun_out/${TAG}_n$N.err > gpurun_out/${TAG}_n$N.json
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 20 --warmup 3 --no-e2e 2> gpurun_out/${TAG}_n$N.err > gpurun_out/${TAG}_n$N.json
  fi
  python - <<PY
import json
for l in open("gpurun_out/${TAG}_n$N.json"):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]; p = d.get("parity") or {}
        print("N=%d  %.3f ms/step  %.1f G p-steps/s  push %.3f ms (frac %.3f)  launches %d  parity ok=%s E=%.6g B=%.6g KE=%.8g migrated=%s  [%s]" % (
            d["n_gpus"], d["ms_per_step"], d["value"] / 1e9, r["ms_per_launch"], r["frac"], d["gpu_launches"], p.get("ok"),
            p.get("e_energy", 0), p.get("b_energy", 0), p.get("kinetic_energy", 0), p.get("migrated"), d["config"]["parallelism"][:60]))
PY
  tail -2 gpurun_out/${TAG}_n$N.err | cut -c1-300
done 2>&1 | tee gpurun_out/${TAG}.log
