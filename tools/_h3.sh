timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 40 --warmup 4 > gpurun_out/r4_bench_n8.log 2>&1
echo rc=$?
grep "^{" gpurun_out/r4_bench_n8.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['value']/1e9, d['ms_per_step'], d.get('host_enqueue_ms_per_step'), d['config']['parallelism'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['gpu_launches'], d['e2e']['value']/1e9, d['clocks'])
"
grep -i "capture\|error\|fail" gpurun_out/r4_bench_n8.log | head
