# round 2, call D (2 GPUs): GPU tests, multi-GPU bit-identity (peer stores + NCCL + render_multi), bench at 1 and 2 GPUs, peer vs NCCL
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py 2>&1 | grep -v "^W\|OMP\|\*\*\*" | tee gpurun_out/dist_check_$N.log
for ex in peer nccl; do
  RTIOW_BENCH_EXCHANGE=$ex timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 30 --warmup 5 --no-other-workloads > gpurun_out/scale_${N}_$ex.json 2> gpurun_out/scale_${N}_$ex.err
  tail -2 gpurun_out/scale_${N}_$ex.err | grep -v "OMP\|\*\*\*"
  python - <<PY
import json
f = "gpurun_out/scale_${N}_$ex.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms", round(d["ms_per_step"], 3), "kernel ms", round(d["roofline"]["kernel_ms_per_launch"], 3), d["frame_crc32"], d["frame_crc32_same_on_all_ranks"], d["config"]["exchange"][:60])
except Exception as e:
    print(f, "n/a", e)
PY
done
