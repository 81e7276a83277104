# round 2, call H: which of the round-2 additions costs the final scene's general kernel its time (code size bisect)
for b in _build _build_xo _build_xb _build_xp _build_xobp; do
  for fuse in 1 0; do
    if [ $fuse = 1 ] && { [ $b = _build_xp ] || [ $b = _build_xobp ]; }; then continue; fi
    echo "== $b fuse=$fuse"
    RTIOW_B200_BUILD_DIR=$b RTIOW_B200_FUSE_PRISMS=$fuse SWEEP_REPS=3 SWEEP_THREADS=0 SWEEP_MODES=2 timeout 200 python scripts/gpu_sweep.py final 2>&1 | cut -c1-120
  done
done
