mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 3 -c 1 -o gpurun_out/prof_fast -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fast.log 2>&1
tail -2 gpurun_out/ncu_fast.log
