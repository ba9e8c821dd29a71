#!/bin/bash
# N=4 and N=8 bench lines with the peer-memory exchange (and N=8 with NCCL forced, for the comparison)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=c14
run() {  # N P2P
  N=$1; P=$2
  CPIC_MGPU_P2P=$P timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800+N+P)) bench.py --gpus $N --steps 20 --warmup 3 --no-e2e 2> gpurun_out/${TAG}_n${N}_p$P.err > gpurun_out/${TAG}_n${N}_p$P.json
  python - <<PY
import json
for l in open("gpurun_out/${TAG}_n${N}_p$P.json"):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]; p = d.get("parity") or {}
        print("N=%d  %.3f ms/step  %.1f G p-steps/s  push %.3f ms (frac %.3f) non-push %.3f  launches %d  parity ok=%s E=%.6g B=%.6g KE=%.8g migrated=%s  [%s]" % (
            d["n_gpus"], d["ms_per_step"], d["value"] / 1e9, r["ms_per_launch"], r["frac"], d["ms_per_step"] - r["ms_per_launch"], d["gpu_launches"], p.get("ok"),
            p.get("e_energy", 0), p.get("b_energy", 0), p.get("kinetic_energy", 0), p.get("migrated"), d["config"]["parallelism"][:150]))
PY
  grep -v "^\*\|OMP_NUM\|^$" gpurun_out/${TAG}_n${N}_p$P.err | tail -2 | cut -c1-300
}
{
NG=$(nvidia-smi -L | wc -l)
[ $NG -ge 4 ] && run 4 1
[ $NG -ge 8 ] && run 8 1
[ $NG -ge 8 ] && run 8 0
} 2>&1 | tee gpurun_out/${TAG}.log
