# round 2, call M: pipelined consecutive renders (two slots): full GPU tests, K-step throughput with and without, one rank's eighth
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for p in 1 0 1 0; do
  RTIOW_B200_PIPELINE=$p timeout 200 python bench.py --steps 30 --warmup 5 --no-other-workloads --no-fast-build --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pipeline $p: value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'kernel', round(d['config']['kernel']['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['frame_crc32'])"
done
python - <<'PY'
# K back-to-back renders of one rank's eighth of C2 (what a rank of an 8-GPU run does), device-timed
import os, sys, torch
sys.path.insert(0, os.getcwd())
import rtiow_rust_b200 as R
from rtiow_rust_b200 import api
nx, ny, ns = 1200, 800, 50
for pipe in ("1", "0", "1", "0"):
    os.environ["RTIOW_B200_PIPELINE"] = pipe
    w, c = R.build_scene("book1", nx, ny)
    out = torch.empty((100, nx, 3), dtype=torch.float32, device="cuda")
    for G in (8, 1):
        o = out if G == 8 else torch.empty((ny, nx, 3), dtype=torch.float32, device="cuda")
        for _ in range(5):
            api.render_rows_device(nx, ny, ns, c, w, o, (0, ny), row_step=G * 4, row_band=4)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 40
        e0.record()
        for _ in range(K):
            api.render_rows_device(nx, ny, ns, c, w, o, (0, ny), row_step=G * 4, row_band=4)
        e1.record(); torch.cuda.synchronize()
        print(f"pipeline {pipe}  G={G}: {e0.elapsed_time(e1) / K:.4f} ms per render", flush=True)
    w.close()
PY
