# round 2, last evidence run on one GPU (the GPU tests passed on this build in the call before): smoke, bench, ncu launch list, ncu --set full on C2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv; nproc
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -2 gpurun_out/bench.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-workloads --no-fast-build > gpurun_out/bench_under_ncu.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 4 -c 1 -o gpurun_out/prof_C2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-workloads --no-fast-build > gpurun_out/ncu_C2.log 2>&1
ls -la gpurun_out | grep -E "prof_|launches|bench"
