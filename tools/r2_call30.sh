#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29730+N)) bench.py --gpus $N --steps 20 --warmup 3 --no-e2e 2> gpurun_out/c30_n$N.err > gpurun_out/c30_n$N.json
python - <<PY
import json
for l in open("gpurun_out/c30_n$N.json"):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]; p = d.get("parity") or {}
        print("N=%d  %.3f ms/step  %.1f G p-steps/s  push %.3f ms (frac %.3f) non-push %.3f  launches %d  parity ok=%s E=%.6g B=%.6g KE=%.8g migrated=%s" % (
            d["n_gpus"], d["ms_per_step"], d["value"] / 1e9, r["ms_per_launch"], r["frac"], d["ms_per_step"] - r["ms_per_launch"], d["gpu_launches"], p.get("ok"),
            p.get("e_energy", 0), p.get("b_energy", 0), p.get("kinetic_energy", 0), p.get("migrated")))
PY
done
} 2>&1 | tee gpurun_out/c30.log
