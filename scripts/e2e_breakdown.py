"""Where the end-to-end step spends its host time: fresh scene upload vs render call (C2)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtiow_rust_b200 as R
from rtiow_rust_b200 import _native as N, api
nx, ny, ns = 1200, 800, 50
world, cam = R.build_scene("book1", nx, ny, use_bvh=True)
out = np.empty((ny, nx, 3), np.float32)
def t(f, n=5):
    f(); ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts), sum(ts) / len(ts)
print("upload_fresh ms (min, mean):", t(lambda: world.upload_fresh(0)))
print("validate ms:", t(lambda: world.validate()))
print("render (host out) ms:", t(lambda: api._check(N.abi().rtiow_b200_render(world.gpu(0), C.byref(cam.rec), nx, ny, ns, 1, out.ctypes.data))))
st = world.stats(0); print("trace_ms", st["trace_ms"], "fold", st["reduce_ms"])
def both():
    world.upload_fresh(0)
    api._check(N.abi().rtiow_b200_render(world.gpu(0), C.byref(cam.rec), nx, ny, ns, 1, out.ctypes.data))
print("fresh + render ms:", t(both))
