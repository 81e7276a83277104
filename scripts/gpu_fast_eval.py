"""R2 figures of a tolerance build (RTIOW_B200_FAST_BUILD_DIR) against the oracle and its kernel time beside the parity build's."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtiow_rust_b200 as R
from rtiow_rust_b200 import api
from oracle import oracle_py as O
NT = os.cpu_count()
for name, bvh, nx, ny, ns in (("book1", True, 400, 200, 50), ("final", False, 160, 160, 64), ("cornell", False, 160, 160, 64)):
    fw, cam = R.build_scene(name, nx, ny, use_bvh=bvh, flavour="fast")
    got = R.par_cast(nx, ny, ns, cam, fw).rgb
    want, _, _ = O.Scene(name, nx, ny, top_level_bvh=bvh).render(ns, nthreads=NT)
    q = lambda a: np.clip((255.99 * np.sqrt(np.maximum(a, 0.0))).astype(np.int64), 0, 255)
    print(name, "mean|d| %.3e" % np.abs(got.astype(np.float64) - want).mean(), "ppm within1 %.5f" % (np.abs(q(got) - q(want)) <= 1).mean(),
          "bit-equal %.3f" % (got.view(np.uint32) == want.view(np.uint32)).mean(), flush=True)
for flavour in ("parity", "fast"):
    w, cam = R.build_scene("book1", 1200, 800, flavour=flavour)
    best = 1e9
    for _ in range(6):
        R.par_cast(1200, 800, 50, cam, w)
        best = min(best, w.stats()["trace_ms"])
    print(flavour, "C2 kernel ms", best, flush=True)
