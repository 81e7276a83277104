# 8 GPUs of one box: C2 at 4 and 8 ranks, then BASELINE.json configs[4] (C5: 4800x3200x500, 7.68 G samples) at 8 ranks
mkdir -p gpurun_out
for n in 4 8; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --workload C5 --steps 3 --warmup 3 > gpurun_out/scale_C5_8.json 2> gpurun_out/scale_C5_8.err
tail -3 gpurun_out/scale_C5_8.err | grep -v "OMP\|\*\*\*"
python - <<PY
import json
for f in ("gpurun_out/scale_4.json", "gpurun_out/scale_8.json", "gpurun_out/scale_C5_8.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms", round(d["ms_per_step"], 3), "kernel ms", round(d["roofline"]["kernel_ms_per_launch"], 3), d["config"]["rows_per_gpu"], d["config"]["kernel"])
    except Exception as e:
        print(f, "n/a", e)
PY
