# round 2, call C: parity + tolerance tests, bench with fast_build entry, C4 after the single-primitive medium path
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -s 2>&1 | grep -E "fast build|passed|failed|Error|error" | tail -15
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err
tail -3 gpurun_out/bench_c.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c.json'))
print('C2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'crc', d['frame_crc32'])
print('fast', json.dumps(d['fast_build'])[:900])
for o in d['other_workloads']: print(o['config'], o['value'], o['ms_per_step'], o['kernel']['ms_per_step'])
PY
