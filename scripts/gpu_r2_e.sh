# round 2, call E: all GPU tests (whole-frame parity), bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -12
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err
tail -3 gpurun_out/bench_e.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_e.json'))
print('C2', d['value'], d['ms_per_step'], 'kernel', d['config']['kernel']['ms_per_step'], 'e2e', d['e2e']['value'], 'crc', d['frame_crc32'], d['e2e'].get('frame_crc32'))
print('fast', d['fast_build']['value'], d['fast_build']['kernel_ms_per_step'])
for o in d['other_workloads']: print(o['config'], o['value'], o['ms_per_step'], o['kernel']['ms_per_step'], o['frame_crc32'])
PY
