#!/usr/bin/env python
"""All five BASELINE.json configurations that fit this box, in one run, as one JSON document:
C1 on the host CPU (1 thread and all threads, the C++ restatement of the reference), C1-C4 on one B200 at
FULL size (kernel time from the library's CUDA events), each GPU frame spot-checked against the oracle on
a few scanline bands.  Usage: python scripts/gpu_results_table.py > gpurun_out/results_table.json"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtiow_rust_b200 as R
from oracle import oracle_py as O

CONFIGS = [("C1", "book1", 400, 200, 10, True), ("C2", "book1", 1200, 800, 50, True),
           ("C3", "cornell", 800, 800, 1000, False), ("C4", "final", 800, 800, 5000, False)]
cores = os.cpu_count() or 1
out = {"host_threads": cores, "cpu": {}, "gpu": {}}
sc = O.Scene("book1", 400, 200, top_level_bvh=True)
for nt in (1, cores):
    sc.render(1, nthreads=nt)
    t0 = time.perf_counter(); sc.render(10, nthreads=nt); dt = time.perf_counter() - t0
    out["cpu"][f"C1_{nt}_threads"] = {"seconds": dt, "msamples_per_s": 400 * 200 * 10 / dt / 1e6,
                                      "what": "C++ restatement of the reference (oracle/), not the Rust binary"}
for key, name, nx, ny, ns, bvh in CONFIGS:
    w, c = R.build_scene(name, nx, ny, use_bvh=bvh)
    R.par_cast(nx, ny, min(ns, 8), c, w)                                  # warm-up
    t0 = time.perf_counter(); img = R.par_cast(nx, ny, ns, c, w).rgb; wall = time.perf_counter() - t0
    st = w.stats()
    # parity spot check: 2 scanline bands of the full-size frame, float for float, at the full sample count
    # (bounded: <= ~2 M oracle samples per config)
    rows = max(1, min(4, 2_000_000 // (nx * ns) // 2))
    osc = O.Scene(name, nx, ny, top_level_bvh=bvh)
    diff = 0; checked = 0
    for r0 in (ny // 3, ny - rows):
        want, _, _ = osc.render(ns, rows=(r0, r0 + rows), nthreads=cores)
        diff += int((img[r0:r0 + rows].view(np.uint32) != want.view(np.uint32)).sum()); checked += want.size
    out["gpu"][key] = {"scene": name, "nx": nx, "ny": ny, "ns": ns, "samples": nx * ny * ns,
                       "kernel_ms": st["trace_ms"], "fold_ms": st["reduce_ms"], "passes": st["passes"],
                       "msamples_per_s_kernel": nx * ny * ns / st["trace_ms"] / 1e3,
                       "host_call_seconds": wall, "msamples_per_s_host_call": nx * ny * ns / wall / 1e6,
                       "segments_per_sample": st["segments"] / st["samples"], "traversal": st["traversal"],
                       "block": st["block"], "regs": st["regs_per_thread"], "scene_in_smem": st["scene_in_smem"],
                       "oracle_floats_checked": checked, "oracle_floats_differing": diff}
    w.close()
print(json.dumps(out, indent=1))
