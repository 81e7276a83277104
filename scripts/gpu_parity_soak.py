#!/usr/bin/env python
"""Whole-frame bit-exactness soak: far more samples than the pytest suite compares (the order-dependent hit that round 1
missed occurred once in 4 M samples).  Prints differing pixels per case; exits non-zero if any."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtiow_rust_b200 as R
from oracle import oracle_py as O

NT = os.cpu_count() or 8
CASES = [("book1", True, 1200, 800, 400), ("cornell", False, 800, 800, 96), ("final", False, 800, 800, 48),
         ("final", True, 800, 800, 24), ("simple_light", True, 800, 800, 24), ("cornell_smoke", False, 800, 800, 32),
         ("kitchen_sink", True, 800, 600, 48), ("book1_head", True, 800, 400, 64)]
bad_total = 0
for name, bvh, nx, ny, ns in CASES:
    world, cam = R.build_scene(name, nx, ny, use_bvh=bvh)
    got = R.par_cast(nx, ny, ns, cam, world).rgb
    t0 = time.time()
    want, _, cnt = O.Scene(name, nx, ny, top_level_bvh=bvh).render(ns, nthreads=NT, want_counters=True)
    bad = int((got.view(np.uint32) != want.view(np.uint32)).any(axis=2).sum())
    segs_equal = world.stats()["segments"] == cnt["segments"]
    print(f"{name} bvh={bvh} {nx}x{ny}x{ns}: {nx * ny * ns / 1e6:.0f} M samples, differing pixels = {bad} of {nx * ny}, "
          f"segment count equal = {segs_equal} (oracle {time.time() - t0:.0f} s on {NT} threads)", flush=True)
    bad_total += bad + (0 if segs_equal else 1)
    world.close()
sys.exit(1 if bad_total else 0)
