# final scene: barrier group size / phase sync / thread count with the lean kernel
for g in 24 12 8 6 4 3; do
  echo "== PHASE_GROUP=$g"; RTIOW_B200_PHASE_GROUP=$g SWEEP_REPS=3 SWEEP_THREADS=0 SWEEP_MODES=2 timeout 200 python scripts/gpu_sweep.py final 2>&1 | cut -c1-110
done
for t in 512 256; do for g in 16 8 4; do
  echo "== threads=$t PHASE_GROUP=$g"; RTIOW_B200_PHASE_GROUP=$g SWEEP_REPS=3 SWEEP_THREADS=$t SWEEP_MODES=2 timeout 200 python scripts/gpu_sweep.py final 2>&1 | cut -c1-110
done; done
echo "== refill lanes"; for r in 1 4 8 16; do RTIOW_B200_REFILL_LANES=$r SWEEP_REPS=3 SWEEP_THREADS=0 SWEEP_MODES=2 timeout 200 python scripts/gpu_sweep.py final 2>&1 | cut -c1-110; done
