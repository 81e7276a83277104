# round 2, call B: parity tests, then C3/C4 kernel time with and without prism fusion / phase barriers, ncu of the C4 kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for f in 1 0; do
  echo "== FUSE_PRISMS=$f"
  RTIOW_B200_FUSE_PRISMS=$f SWEEP_REPS=3 SWEEP_THREADS=0 SWEEP_MODES=0 timeout 300 python scripts/gpu_sweep.py cornell final final_bvh 2>&1 | cut -c1-260
done
for ps in 0 1; do
  echo "== fused, PHASE_SYNC=$ps"
  RTIOW_B200_PHASE_SYNC=$ps SWEEP_REPS=3 SWEEP_THREADS=0,512 SWEEP_MODES=0,2 timeout 300 python scripts/gpu_sweep.py final 2>&1 | cut -c1-260
done
NS=50 bash scripts/gpu_ncu_scene.sh C4
ls -la gpurun_out | tail -5
