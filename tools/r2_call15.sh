#!/bin/bash
# 4 GPUs: (a) the native stepper against the oracle on 4 ranks, with and without a rank held back before every send (race
# detector for the peer-memory landing buffers), (b) the TMA bulk-store variant of k_push3: parity + timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== test_mgpu on 4 ranks, peer memory, no skew"
CPIC_TEST_WORLD=4 CPIC_REQUIRE_P2P=1 CPIC_P2P_TIMEOUT_S=20 timeout 400 python -m pytest tests/test_mgpu.py -x -q -m gpu -k "slab" 2>&1 | tail -2
for SK in 200 2000; do
echo "== test_mgpu on 4 ranks, peer memory, rank 1 held back $SK us before every send"
CPIC_TEST_WORLD=4 CPIC_REQUIRE_P2P=1 CPIC_P2P_TIMEOUT_S=20 CPIC_P2P_SKEW_US=$SK timeout 400 python -m pytest tests/test_mgpu.py -x -q -m gpu -k "slab" 2>&1 | tail -2
done
echo "== bench N=4, skew 300 us (parity block must equal the other lines: E=143.437 B=7.2685 migrated=17258365)"
for SK in 0 300; do
CPIC_P2P_SKEW_US=$SK timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2985$((SK/300)) bench.py --gpus 4 --steps 20 --warmup 3 --no-e2e 2> gpurun_out/c15_n4_s$SK.err > gpurun_out/c15_n4_s$SK.json
python - <<PY
import json
for l in open("gpurun_out/c15_n4_s$SK.json"):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]; p = d.get("parity") or {}
        print("N=%d  %.3f ms/step  push %.3f ms non-push %.3f  parity ok=%s E=%.6g B=%.6g KE=%.8g migrated=%s" % (
            d["n_gpus"], d["ms_per_step"], r["ms_per_launch"], d["ms_per_step"] - r["ms_per_launch"], p.get("ok"), p.get("e_energy", 0), p.get("b_energy", 0), p.get("kinetic_energy", 0), p.get("migrated")))
PY
done
echo "== bulk-store variant: parity (k_push3 forced on small grids)"
CPIC_LIB=$PWD/cabanapic_b200/libcabanapic_b200_bulk.so CPIC_PUSH2_PRIV=0 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "push_reorder or sorted_steps or fused_step or slab or reflect or species" 2>&1 | tail -3
echo "== timing: product, bulk"
bash tools/r2_vars.sh bulk
CPIC_LIB=$PWD/cabanapic_b200/libcabanapic_b200_bulk.so timeout 300 python tools/probe_reorder.py 256 256 256 64 6 reorder | tail -1
} 2>&1 | tee gpurun_out/c15.log
