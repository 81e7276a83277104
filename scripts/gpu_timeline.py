"""Timeline of K back-to-back renders of one rank's 1/G of C2 (library events, RTIOW_B200_TIMELINE=1; printed at scene_destroy)."""
import os, sys, torch
os.environ["RTIOW_B200_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtiow_rust_b200 as R
from rtiow_rust_b200 import api
nx, ny, ns = 1200, 800, 50
G = int(os.environ.get("SHARE_G", "8"))
w, c = R.build_scene("book1", nx, ny)
o = torch.empty((ny // G + 8, nx, 3), dtype=torch.float32, device="cuda")
for _ in range(12):
    api.render_rows_device(nx, ny, ns, c, w, o, (0, ny), row_step=G * 4, row_band=4)
torch.cuda.synchronize()
w.close()
