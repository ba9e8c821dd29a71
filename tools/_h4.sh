# (1) launch list of the bench command (one metric, serialised; the numbers it prints are not bench values)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r03_launches_bench_c5.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r4_ncu_launch.log 2>&1
echo rc1=$?
# (2) full capture of one steady-state reordering push at 256x256x64 x 64 ppc (4th k_push2 launch)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_push2 -s 3 -c 1 -o gpurun_out/r03_push2_steady_256x256x64 -f python tools/probe_reorder.py 256 256 64 64 5 reorder > gpurun_out/r4_ncu_full.log 2>&1
echo rc2=$?
ls -la gpurun_out/ | tail -5
