# 8 GPUs of one box, short: bit-identity at 8 ranks, C2 at 8 / 4 / 2 / 1 ranks (same box, so the ratios are clean)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py 2>&1 | grep -v "^W\|OMP\|\*\*\*" | tee gpurun_out/dist_check_8.log
for n in 8 4 2 1; do
  if [ $n = 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-other-workloads --no-fast-build --no-cpu-baseline > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 30 --warmup 5 --no-other-workloads > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  fi
done
python - <<PY
import json
for f in ("scale_1", "scale_2", "scale_4", "scale_8"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms", round(d["ms_per_step"], 3), "kernel ms", round(d["roofline"]["kernel_ms_per_launch"], 3),
              d["frame_crc32"], d["frame_crc32_same_on_all_ranks"], d["config"]["exchange"][:40])
    except Exception as e:
        print(f, "n/a", e)
PY
