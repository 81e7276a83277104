# N GPUs of one box: NCCL correctness of the sharded path, then the bench at 1..N
set -x
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py 2>&1 | grep -v "^W\|OMP" | tee gpurun_out/dist_check_$N.log
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  if [ $n -eq 1 ]; then python bench.py --steps 20 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
       python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $n --steps 20 --warmup 3 --contiguous > gpurun_out/scale_${n}_contiguous.json 2> gpurun_out/scale_${n}_contiguous.err
  fi
  tail -2 gpurun_out/scale_$n.err
  python - <<PY
import json
for f in ("gpurun_out/scale_$n.json", "gpurun_out/scale_${n}_contiguous.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms", round(d["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 3), d["config"]["rows_per_gpu"])
    except Exception as e:
        print(f, "n/a", e)
PY
done
