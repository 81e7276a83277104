# round 2, call L: what costs one rank's eighth of C2 its 13 % (kernel 1.53 ms against T(1)/8 = 1.33 ms)?
echo "== default"; SHARD_G=1,8 timeout 100 python scripts/gpu_shard_time.py
for v in "RTIOW_B200_REFILL_LANES=4" "RTIOW_B200_REFILL_LANES=8" "RTIOW_B200_REFILL_LANES=20" "RTIOW_B200_UNIT_ORDER=0" "RTIOW_B200_SAMPLE_CHUNK=2" "RTIOW_B200_SAMPLE_CHUNK=4" "RTIOW_B200_CTA_THREADS=512" "RTIOW_B200_CTA_THREADS=768"; do
  echo "== $v"; env $v SHARD_G=8 timeout 100 python scripts/gpu_shard_time.py
done
echo "== book1/cornell/final kernel times"
SWEEP_REPS=4 SWEEP_THREADS=0 SWEEP_MODES=0 timeout 300 python scripts/gpu_sweep.py book1 cornell final 2>&1 | cut -c1-150
