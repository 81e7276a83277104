# same-box A/B of two in-tree builds: usage gpu_ab_builds.sh <dirA> <dirB> [scenes...]
A=$1; B=$2; shift 2; SC=${@:-book1}
for b in $A $B $A $B; do
  echo "== $b"
  RTIOW_B200_BUILD_DIR=$b SWEEP_REPS=4 SWEEP_THREADS=0 SWEEP_MODES=0 timeout 300 python scripts/gpu_sweep.py $SC 2>&1 | cut -c1-150
done
