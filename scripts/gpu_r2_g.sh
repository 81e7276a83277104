# round 2, call G: same-box A/B against the round-1 library (r01_snapshot/), unit order, thread counts
echo "== round-1 library"
(cd r01_snapshot && SWEEP_REPS=3 SWEEP_THREADS=0 SWEEP_MODES=0,2 timeout 300 python scripts/gpu_sweep.py book1 cornell final 2>&1 | cut -c1-200)
echo "== current, top-first units"
RTIOW_B200_UNIT_ORDER=0 SWEEP_REPS=3 SWEEP_THREADS=0 SWEEP_MODES=0,2 timeout 300 python scripts/gpu_sweep.py book1 cornell final 2>&1 | cut -c1-200
echo "== current, bottom-first units"
SWEEP_REPS=3 SWEEP_THREADS=0 SWEEP_MODES=2 timeout 300 python scripts/gpu_sweep.py final 2>&1 | cut -c1-200
echo "== current, unfused, top-first"
RTIOW_B200_FUSE_PRISMS=0 RTIOW_B200_UNIT_ORDER=0 SWEEP_REPS=3 SWEEP_THREADS=0 SWEEP_MODES=2 timeout 300 python scripts/gpu_sweep.py final 2>&1 | cut -c1-200
