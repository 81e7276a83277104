#!/usr/bin/env python
"""Final scene (C4 shape, 100 spp): exact nodes in shared memory vs conservative nodes read from global memory."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtiow_rust_b200 as R
nx, ny, ns = 800, 800, 100
w, c = R.build_scene("final", nx, ny, use_bvh=False)
for trav, fg, thr in ((0, False, 512), (0, True, 512), (0, True, 768), (0, True, 256), (2, False, 512), (2, True, 512)):
    w.set_traversal(trav)
    w.set_tuning(cta_threads=thr, force_global=fg)
    best = 1e9
    for _ in range(3):
        R.par_cast(nx, ny, ns, c, w)
        st = w.stats()
        best = min(best, st["trace_ms"])
    print(f"final trav_req={trav} force_global={fg} thr={thr}: trav={st['traversal']} in_smem={st['scene_in_smem']} scene_bytes={st['scene_bytes']} "
          f"regs={st['regs_per_thread']} trace {best:.2f} ms -> {st['samples'] / best / 1e3:.1f} Msamples/s", flush=True)
