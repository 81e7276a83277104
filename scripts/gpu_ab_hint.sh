# same-box A/B: the build before the tile-order hint (_build_head), the new build with the hint off and on
exec > gpurun_out/ab_hint.log 2>&1
for rep in 1 2; do
  echo "== _build_head"; RTIOW_B200_BUILD_DIR=_build_head SHARE_GS=8,4,1 python scripts/gpu_pipelined_share.py
  echo "== _build hint off"; RTIOW_B200_ORDER_HINT=0 SHARE_GS=8,4,1 python scripts/gpu_pipelined_share.py
  echo "== _build hint on"; SHARE_GS=8,4,1 python scripts/gpu_pipelined_share.py
done
