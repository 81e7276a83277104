mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
SWEEP_REPS=2 SWEEP_THREADS=768,512 SWEEP_MODES=0,2 SWEEP_SCHED=1:0:0,2:2:24 timeout 120 python scripts/gpu_sweep.py book1 cornell final 2>&1 | cut -c1-175 | tee gpurun_out/sweep_v6a.log
RTIOW_B200_FORCE_GLOBAL=1 SWEEP_REPS=2 SWEEP_THREADS=512 SWEEP_MODES=0 SWEEP_SCHED=1:0:0 timeout 120 python scripts/gpu_sweep.py final 2>&1 | cut -c1-175 | tee -a gpurun_out/sweep_v6a.log
