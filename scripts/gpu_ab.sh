# same-box A/B of two in-tree builds (make OUT=_build_a / default _build): interleaved, three rounds
for r in 1 2 3; do for d in _build_a _build; do
echo "== $d"; RTIOW_B200_BUILD_DIR=$d SWEEP_REPS=3 SWEEP_THREADS=${THREADS:-0} SWEEP_MODES=0 timeout 100 python scripts/gpu_sweep.py ${SCENES:-book1} 2>&1 | cut -c1-110
done; done
