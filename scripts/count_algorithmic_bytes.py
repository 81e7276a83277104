#!/usr/bin/env python
"""Regenerates profiles/algorithmic_bytes.json: SURVEY §8(d)'s bytes per sample, counted by the
oracle on the REFERENCE's own traversal (left-first median-split BVH / linear list) at the BASELINE
configurations.  C3/C4 are counted at 16 spp (the per-sample averages do not depend on ns beyond
noise; the full 1000/5000 spp frames would take hours on the CPU).

  python scripts/count_algorithmic_bytes.py            # all four configs (~1 min on 8 threads)
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O  # noqa: E402  (test/measurement infrastructure)

CONFIGS = {
    "C1": dict(scene="book1", nx=400, ny=200, ns=10, top_level_bvh=True),
    "C2": dict(scene="book1", nx=1200, ny=800, ns=50, top_level_bvh=True),
    "C3": dict(scene="cornell", nx=800, ny=800, ns=16, top_level_bvh=False),
    "C4": dict(scene="final", nx=800, ny=800, ns=16, top_level_bvh=False),
}


def main():
    nt = os.cpu_count() or 1
    out = {}
    for key, c in CONFIGS.items():
        sc = O.Scene(c["scene"], c["nx"], c["ny"], top_level_bvh=c["top_level_bvh"])
        t0 = time.time()
        _, _, cnt = sc.render(c["ns"], nthreads=nt, want_counters=True)
        dt = time.time() - t0
        n = cnt["samples"]
        rec = dict(c)
        rec.update(samples=n, segments_per_sample=cnt["segments"] / n, node_tests_per_sample=cnt["node_tests"] / n,
                   sphere_tests_per_sample=cnt["sphere_tests"] / n, rect_tests_per_sample=cnt["rect_tests"] / n,
                   medium_evals_per_sample=cnt["medium_evals"] / n, draws_per_sample=cnt["draws"] / n,
                   max_segments=cnt["max_segments"],
                   bytes_per_sample=O.algorithmic_bytes_per_sample(cnt, c["ns"]))
        rec[f"oracle_seconds_{nt}_threads"] = dt
        out[key] = rec
        print(key, json.dumps(rec), flush=True)
    with open(os.path.join(ROOT, "profiles", "algorithmic_bytes.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
