# round 2, final evidence run on one GPU: smoke, all GPU tests, both bench arms, ncu launch list, ncu --set full on C2 / C3 / C4
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv; nproc
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-workloads --no-fast-build > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 3 -c 1 -o gpurun_out/prof_C2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-workloads --no-fast-build > gpurun_out/ncu_C2.log 2>&1
NS=100 bash scripts/gpu_ncu_scene.sh C3
NS=50 bash scripts/gpu_ncu_scene.sh C4
ls -la gpurun_out | grep -E "prof_|launches|bench"
