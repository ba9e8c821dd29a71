#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench (default)"
timeout 900 python bench.py > gpurun_out/verify_bench_n1.json 2> gpurun_out/verify_bench.err; tail -2 gpurun_out/verify_bench.err | cut -c1-300
python - <<PY
import json
for l in open("gpurun_out/verify_bench_n1.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")}))
        print("roofline", json.dumps(d["roofline"])[:600])
        print("e2e", json.dumps(d.get("e2e"))[:300])
        print("cpu", json.dumps(d.get("cpu_baseline"))[:300])
        print("extra", json.dumps(d.get("extra"))[:1800])
PY
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-600
} 2>&1 | tee gpurun_out/verify.log
