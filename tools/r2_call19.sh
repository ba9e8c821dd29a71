#!/bin/bash
# 2 GPUs: the streamed host step of the multi-GPU layer -- parity, then bench N=2 with e2e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== test_mgpu (2 ranks)"
CPIC_REQUIRE_P2P=1 CPIC_P2P_TIMEOUT_S=20 timeout 600 python -m pytest tests/test_mgpu.py -x -q -m gpu 2>&1 | tail -15
echo "== test_mgpu step_host, 1 rank (open z)"
CUDA_VISIBLE_DEVICES=0 timeout 300 python -m pytest tests/test_mgpu.py -x -q -m gpu -k "step_host" 2>&1 | tail -3
echo "== new single-GPU tests"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_facade.py -x -q -m gpu -k "odd_count or two_stream_short or dioctron_full or step_host" 2>&1 | tail -3
echo "== bench N=2 with e2e (256^3)"
CPIC_P2P_TIMEOUT_S=120 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/c19_n2.err > gpurun_out/c19_n2.json
python - <<PY
import json
for l in open("gpurun_out/c19_n2.json"):
    if l.startswith("{"):
        d = json.loads(l); p = d.get("parity") or {}
        print("N=%d  %.3f ms/step  parity ok=%s E=%.6g B=%.6g migrated=%s" % (d["n_gpus"], d["ms_per_step"], p.get("ok"), p.get("e_energy", 0), p.get("b_energy", 0), p.get("migrated")))
        print("e2e", json.dumps(d.get("e2e"))[:700])
PY
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/c19_n2.err | tail -3 | cut -c1-300
} 2>&1 | tee gpurun_out/c19.log
