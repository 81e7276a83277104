# round 2, call I: lean general kernel
timeout 900 python -m pytest tests -x -q -m gpu -k "specialis or golden or every_scene or whole or execution_shape" 2>&1 | tail -4
echo "== lean (default)"
SWEEP_REPS=3 SWEEP_THREADS=0,512 SWEEP_MODES=0,2 timeout 400 python scripts/gpu_sweep.py final final_bvh 2>&1 | cut -c1-200
echo "== general kernel (RTIOW_B200_SPECIALISE=0)"
RTIOW_B200_SPECIALISE=0 SWEEP_REPS=3 SWEEP_THREADS=0 SWEEP_MODES=2 timeout 400 python scripts/gpu_sweep.py final 2>&1 | cut -c1-200
echo "== lean, phase sync variants"
for ps in 0 1; do RTIOW_B200_PHASE_SYNC=$ps SWEEP_REPS=3 SWEEP_THREADS=0 SWEEP_MODES=2 timeout 200 python scripts/gpu_sweep.py final 2>&1 | cut -c1-200; done
