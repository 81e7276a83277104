#!/usr/bin/env python
"""One GPU doing one rank's share of C2: kernel time of rows rank, rank + G, ... for G = 1, 2, 4, 8 against T(1)/G."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtiow_rust_b200 as R
from rtiow_rust_b200 import api
nx, ny, ns = 1200, 800, 50
w, c = R.build_scene("book1", nx, ny, use_bvh=True)
out = torch.empty((ny, nx, 3), dtype=torch.float32, device="cuda")
t1 = None
GS = [int(x) for x in os.environ.get("SHARD_G", "1,2,4,8").split(",")]
for G, B in [(g, 4) for g in GS]:
    for rank in sorted({0, G - 1}):
        best = 1e9
        for _ in range(4):
            api.render_rows_device(nx, ny, ns, c, w, out, (rank * B, ny), row_step=G * B, row_band=B)
            torch.cuda.synchronize()
            best = min(best, w.stats()["trace_ms"])
        t1 = t1 or best * G
        print(f"G={G} band={B} rank={rank}: trace {best:.3f} ms, ideal {t1 / G:.3f} ms, efficiency {t1 / G / best:.3f}", flush=True)
