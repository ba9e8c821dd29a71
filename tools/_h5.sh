timeout 100 python tools/probe_reorder.py 256 256 64 64 4 reorder > gpurun_out/r4_probe2.log 2>&1
cat gpurun_out/r4_probe2.log
timeout 200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
