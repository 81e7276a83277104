# round 2, call K: 64-byte conservative-test nodes + two-size work units (new _build) against round-2-so-far (_build_n112), same box
timeout 900 python -m pytest tests -x -q -m gpu -k "whole or golden or every_scene or traversal or specialis or execution_shape or ragged or row_blocks or strided" 2>&1 | tail -3
for b in _build_n112 _build _build_n112 _build; do
  echo "== $b"
  RTIOW_B200_BUILD_DIR=$b SWEEP_REPS=4 SWEEP_THREADS=0 SWEEP_MODES=0 timeout 300 python scripts/gpu_sweep.py book1 cornell 2>&1 | cut -c1-150
done
echo "== _build, one unit size (8) as before"
RTIOW_B200_SAMPLE_CHUNK=8 SWEEP_REPS=4 SWEEP_THREADS=0 SWEEP_MODES=0 timeout 300 python scripts/gpu_sweep.py book1 2>&1 | cut -c1-150
echo "== final: conservative 64-byte-node tree (mode 0, now fits the L1 rule) vs exact tree (mode 2)"
SWEEP_REPS=3 SWEEP_THREADS=0,512 SWEEP_MODES=0,2 timeout 300 python scripts/gpu_sweep.py final final_bvh 2>&1 | cut -c1-200
echo "== one GPU doing one rank's share of C2"
for b in _build_n112 _build; do echo "-- $b"; RTIOW_B200_BUILD_DIR=$b timeout 200 python scripts/gpu_shard_time.py; done
