# N GPUs of one box (default 2): GPU parity tests, NCCL correctness of the sharded path, then the bench at 2..N
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py 2>&1 | grep -v "^W\|OMP\|\*\*\*" | tee gpurun_out/dist_check_$N.log
for n in 2 4 8; do
  [ $n -le $N ] || continue
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  tail -2 gpurun_out/scale_$n.err | grep -v "OMP\|\*\*\*"
  python - <<PY
import json
f = "gpurun_out/scale_$n.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms", round(d["ms_per_step"], 3), "kernel ms", round(d["roofline"]["kernel_ms_per_launch"], 3), d["config"]["rows_per_gpu"])
except Exception as e:
    print(f, "n/a", e)
PY
done
