import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")


def golden_cases():
    z = np.load(GOLDEN)
    for key in z.files:
        if key.endswith("|segments"):
            continue
        name, nx, ny, ns, bvh = key.split("|")
        yield key, name, int(nx), int(ny), int(ns), bool(int(bvh)), z[key], int(z[key + "|segments"][0])


def bits_equal(a, b):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def n_diff(a, b):
    return int((np.ascontiguousarray(a, np.float32).view(np.uint32) != np.ascontiguousarray(b, np.float32).view(np.uint32)).sum())
