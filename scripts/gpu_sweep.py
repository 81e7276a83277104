#!/usr/bin/env python
"""Kernel-time sweep over execution shapes (threads per CTA, CTAs per SM, traversal mode) for the
BASELINE scenes.  Prints one line per point; timing is the library's own CUDA events around the
megakernel.  Usage: python scripts/gpu_sweep.py [book1|cornell|final ...]"""
import itertools
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtiow_rust_b200 as R  # noqa: E402

CASES = {"book1": ("book1", 1200, 800, 50, True), "cornell": ("cornell", 800, 800, 100, False),
         "final": ("final", 800, 800, 100, False), "final_bvh": ("final", 800, 800, 100, True)}
names = sys.argv[1:] or ["book1", "cornell", "final"]
threads = [int(x) for x in os.environ.get("SWEEP_THREADS", "0,512,768").split(",")]
cps = [int(x) for x in os.environ.get("SWEEP_CPS", "0").split(",")]
modes = [int(x) for x in os.environ.get("SWEEP_MODES", "0,2,1").split(",")]
reps = int(os.environ.get("SWEEP_REPS", "3"))

for key in names:
    name, nx, ny, ns, bvh = CASES[key]
    w, c = R.build_scene(name, nx, ny, use_bvh=bvh)
    for mode, thr, cp in itertools.product(modes, threads, cps):
        w.set_traversal(mode)
        w.set_tuning(cta_threads=thr, ctas_per_sm=cp)
        best = None
        for _ in range(reps):
            t = time.time()
            try:
                R.par_cast(nx, ny, ns, c, w)
            except R.RtiowError as e:
                print(f"{key} mode={mode} thr={thr}: {e}")
                break
            dt = time.time() - t
            st = w.stats()
            if best is None or st["trace_ms"] < best[0]:
                best = (st["trace_ms"], st["reduce_ms"], dt)
        if best is None:
            continue
        print(f"{key} {nx}x{ny}x{ns} mode={mode} thr={thr} cps={cp} trav={st['traversal']}: trace {best[0]:.2f} ms fold {best[1]:.2f} ms -> "
              f"{st['samples'] / best[0] / 1e3:.1f} Msamples/s (kernel), e2e {best[2] * 1e3:.1f} ms, grid {st['grid']} regs "
              f"{st['regs_per_thread']} smem {st['dyn_smem_bytes']} in_smem {st['scene_in_smem']} nodes {st['accel_nodes']} "
              f"segs/sample {st['segments'] / st['samples']:.3f}", flush=True)
