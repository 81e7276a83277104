"""Generates tests/golden/golden_v1.npz from the CPU oracle.

The reference (Rust) cannot be built or run in this environment and holds no golden vectors for
the render path, so these fixtures are the ORACLE's outputs ("parity unpinned"): they
pin the oracle against accidental change and give the GPU tests a target that does not need the
oracle to be rebuilt.  Re-run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O  # noqa: E402

CASES = [  # name, nx, ny, ns, top_level_bvh
    ("book1", 40, 20, 4, True), ("book1", 16, 8, 2, False), ("book1_head", 32, 16, 4, True),
    ("cornell", 24, 24, 6, False), ("cornell", 24, 24, 6, True), ("bench_cornell", 10, 10, 4, True),
    ("final", 24, 24, 6, False), ("final", 24, 24, 6, True), ("motion_test", 24, 24, 4, False),
    ("volume_test", 24, 24, 4, False), ("simple_light", 24, 24, 3, True), ("kitchen_sink", 32, 24, 6, False),
    ("kitchen_sink", 32, 24, 6, True),
    # round 2: ConstantMedium with rect_prism and Bvh boundaries (object.rs:533-541 takes any Object)
    ("cornell_smoke", 32, 32, 6, False), ("cornell_smoke", 32, 32, 6, True),
]
SEED = 0xDEADBEEF


def main():
    out = {}
    for name, nx, ny, ns, bvh in CASES:
        img, _, cnt = O.Scene(name, nx, ny, top_level_bvh=bvh).render(ns, seed=SEED, nthreads=4, want_counters=True)
        key = f"{name}|{nx}|{ny}|{ns}|{int(bvh)}"
        out[key] = img
        out[key + "|segments"] = np.array([cnt["segments"]], np.uint64)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    if os.path.exists(path):   # fixtures are append-only: an existing case must come out bit-identical
        old = np.load(path)
        for k in old.files:
            assert k in out and np.array_equal(old[k].view(np.uint8), out[k].view(np.uint8)), f"golden case {k} changed"
    np.savez_compressed(path, **out)
    print(f"wrote {len(CASES)} cases")


if __name__ == "__main__":
    main()
