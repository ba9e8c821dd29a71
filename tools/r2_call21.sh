#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
nvidia-smi topo -m 2>&1 | head -12 | cut -c1-160
N=8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29718 bench.py --gpus $N --steps 6 --warmup 3 --e2e-steps 3 2> gpurun_out/c21_n$N.err > gpurun_out/c21_n$N.json
python - <<PY
import json
for l in open("gpurun_out/c21_n8.json"):
    if l.startswith("{"):
        d = json.loads(l); p = d.get("parity") or {}
        print("N=%d  %.3f ms/step  parity ok=%s" % (d["n_gpus"], d["ms_per_step"], p.get("ok")))
        print("e2e", json.dumps(d.get("e2e"))[:200], (d.get("e2e") or {}).get("host_buffers_numa_local"))
PY
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/c21_n8.err | tail -3 | cut -c1-300
} 2>&1 | tee gpurun_out/c21.log
