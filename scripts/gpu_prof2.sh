set -x
mkdir -p gpurun_out
SWEEP_THREADS=512,768 SWEEP_MODES=0 SWEEP_SCHED=1 python scripts/gpu_sweep.py book1 final 2>&1 | tee gpurun_out/sweep_v3b.log
RTIOW_B200_SCHEDULE=1 RTIOW_B200_CTA_THREADS=512 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 3 -c 1 -o gpurun_out/prof_lockstep -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
RTIOW_B200_SCHEDULE=0 RTIOW_B200_THR_SLOW=24 RTIOW_B200_THR_LEAF=8 RTIOW_B200_CTA_THREADS=512 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 3 -c 1 -o gpurun_out/prof_interleaved -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
cp rtiow-rust_b200/_build/librtiow_b200.so gpurun_out/librtiow_b200.profiled.so
