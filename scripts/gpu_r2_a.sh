# round 2, call A: smoke + parity tests + bench (both arms) + unit-order A/B
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
nproc
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
tail -3 gpurun_out/bench_a.err
cat gpurun_out/bench_a.json
for o in 0 1 0 1; do
  RTIOW_B200_UNIT_ORDER=$o timeout 120 python bench.py --steps 20 --warmup 3 --no-other-workloads --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('unit_order $o', d['value'], d['ms_per_step'], d['config']['kernel']['ms_per_step'], d['e2e']['value'], d['frame_crc32'])"
done
