# round 2, call F: parity subset, then kernel-time sweeps: fast (conservative) tree now fits the final scene; prism variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "whole or golden or every_scene or traversal or specialis" 2>&1 | tail -4
echo "== default build"
SWEEP_REPS=3 SWEEP_THREADS=0,512 SWEEP_MODES=0,2 timeout 400 python scripts/gpu_sweep.py book1 cornell final final_bvh 2>&1 | cut -c1-230
echo "== FUSE_PRISMS=0"
RTIOW_B200_FUSE_PRISMS=0 SWEEP_REPS=3 SWEEP_THREADS=0 SWEEP_MODES=0,2 timeout 300 python scripts/gpu_sweep.py cornell final 2>&1 | cut -c1-230
echo "== prism inline build"
RTIOW_B200_BUILD_DIR=_build_pi SWEEP_REPS=3 SWEEP_THREADS=0 SWEEP_MODES=0,2 timeout 300 python scripts/gpu_sweep.py cornell final 2>&1 | cut -c1-230
echo "== phase sync variants (fast tree)"
for ps in 0 1; do RTIOW_B200_PHASE_SYNC=$ps SWEEP_REPS=3 SWEEP_THREADS=0 SWEEP_MODES=0 timeout 200 python scripts/gpu_sweep.py final 2>&1 | cut -c1-230; done
