mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
SWEEP_REPS=3 SWEEP_THREADS=768,512 SWEEP_MODES=0 timeout 120 python scripts/gpu_sweep.py book1 cornell final 2>&1 | cut -c1-200 | tee gpurun_out/sweep_quick.log
timeout 100 python scripts/e2e_breakdown.py
