"""K back-to-back renders of one rank's 1/G of C2 (what a rank of a G-GPU run does), device-timed.  env SHARE_THREADS, SHARE_GS."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtiow_rust_b200 as R
from rtiow_rust_b200 import api
nx, ny, ns = 1200, 800, 50
thr = int(os.environ.get("SHARE_THREADS", "0"))
w, c = R.build_scene("book1", nx, ny)
w.set_tuning(cta_threads=thr)
for G in [int(g) for g in os.environ.get("SHARE_GS", "8,4,1").split(",")]:
    o = torch.empty((ny // G + 8, nx, 3), dtype=torch.float32, device="cuda")
    for _ in range(5):
        api.render_rows_device(nx, ny, ns, c, w, o, (0, ny), row_step=G * 4, row_band=4)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 40
    e0.record()
    for _ in range(K):
        api.render_rows_device(nx, ny, ns, c, w, o, (0, ny), row_step=G * 4, row_band=4)
    e1.record(); torch.cuda.synchronize()
    print(f"threads {thr} G={G}: {e0.elapsed_time(e1) / K:.4f} ms per render", flush=True)
