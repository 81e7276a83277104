# usage: gpu_ncu_scene.sh <workload C3|C4> : one full ncu capture of the megakernel on another BASELINE scene (reduced spp via env)
mkdir -p gpurun_out
W=${1:-C4}
RTIOW_BENCH_NS=${NS:-50} timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -o gpurun_out/prof_$W -f python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline --no-other-workloads --no-fast-build > gpurun_out/ncu_$W.log 2>&1
tail -2 gpurun_out/ncu_$W.log
