#!/bin/bash
# N=8 bench line of the final code (peer-memory exchange, streamed e2e)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
N=8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29708 bench.py --gpus $N --steps 20 --warmup 3 2> gpurun_out/c20_n$N.err > gpurun_out/c20_n$N.json
python - <<PY
import json
for l in open("gpurun_out/c20_n8.json"):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]; p = d.get("parity") or {}
        print("N=%d  %.3f ms/step  %.1f G p-steps/s  push %.3f ms (frac %.3f) non-push %.3f  launches %d  parity ok=%s E=%.6g B=%.6g KE=%.8g migrated=%s" % (
            d["n_gpus"], d["ms_per_step"], d["value"] / 1e9, r["ms_per_launch"], r["frac"], d["ms_per_step"] - r["ms_per_launch"], d["gpu_launches"], p.get("ok"),
            p.get("e_energy", 0), p.get("b_energy", 0), p.get("kinetic_energy", 0), p.get("migrated")))
        print("e2e", json.dumps(d.get("e2e"))[:400])
PY
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/c20_n8.err | tail -3 | cut -c1-300
} 2>&1 | tee gpurun_out/c20.log
