"""TEST INFRASTRUCTURE: builds and binds tests/kernel_host_harness.cpp (the megakernel's per-path
code compiled for the host).  Never imported by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libkernel_host_harness.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "kernel_host_harness.cpp")
        deps = [src] + [os.path.join(_HERE, "..", "rtiow-rust_b200", "csrc", d, f) for d, f in
                        (("device", "path_logic.cuh"), ("device", "rt_math.cuh"), ("abi", "scene_blob.hpp"), ("abi", "accel_build.hpp"),
                         ("abi", "unit_plan.hpp"))]
        if not os.path.exists(_SO) or any(os.path.getmtime(d) > os.path.getmtime(_SO) for d in deps):
            os.makedirs(os.path.dirname(_SO), exist_ok=True)
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fvisibility=hidden", "-Wl,-Bsymbolic",
                                   "-shared", "-o", _SO, src])
        _lib = C.CDLL(_SO)
        _lib.harness_render.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64,
                                        C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_uint32, C.c_uint32]
        _lib.harness_trace_rays.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
        _lib.harness_unit_coverage.argtypes = [C.c_uint32] * 6 + [C.c_int, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.harness_unit_coverage.restype = C.c_long
    return _lib


def trace_rays(world, rays, accel, fuse_prisms=True):
    """hit_top for explicit rays [n, 7] = (origin, direction, time) -> uint32 [n, 2] = (winner id, bits of t)."""
    rays = np.ascontiguousarray(rays, np.float32)
    out = np.zeros((rays.shape[0], 2), np.uint32)
    rc = lib().harness_trace_rays(C.cast(world.desc, C.c_void_p), rays.shape[0], rays.ctypes.data, out.ctypes.data, int(accel) | (0 if fuse_prisms else 0x200))
    if rc:
        raise RuntimeError(f"harness_trace_rays failed: {rc}")
    return out


def render(world, camera, nx, ny, ns, seed=0xDEADBEEF, rows=None, want_samples=False, accel=True, layout=None, row_step=1,
           row_band=1, lean=False, fuse_prisms=True):
    """rows=(begin, end): bands of row_band rows starting at begin, begin + row_step, ... clipped to end."""
    r0, r1 = rows if rows is not None else (0, ny)
    n_rows = sum(min(row_band, r1 - b) for b in range(r0, r1, row_step))
    img = np.zeros((n_rows, nx, 3), np.float32)
    smp = np.zeros((n_rows, nx, ns, 4), np.float32) if want_samples else None
    rc = lib().harness_render(C.cast(world.desc, C.c_void_p), C.byref(camera.rec), nx, ny, ns, seed, r0, r1,
                              img.ctypes.data, smp.ctypes.data if want_samples else None, int(accel) | (0x100 if lean else 0) | (0 if fuse_prisms else 0x200),
                              layout.ctypes.data if layout is not None else None, row_step, row_band)
    if rc:
        raise RuntimeError(f"harness_render failed: {rc}")
    return img, smp


def unit_coverage(nx, n_rows, s_count, tile_first=0, tile_step=1, forced_chunk=0, open_scene=True, resident_warps=4736, max_strips=4096,
                  order=None):
    """Walks every work unit of one launch like the megakernel's refill does, with the host's own plan.  Returns
    (counts [n_rows, nx, s_count] uint8 = how often each pixel-sample is handed out, staging slots the fold would misplace,
    plan dict)."""
    counts = np.zeros((n_rows, nx, s_count), np.uint8)
    plan = np.zeros(7, np.uint32)
    o = None if order is None else np.ascontiguousarray(order, np.uint32)
    bad = lib().harness_unit_coverage(nx, n_rows, tile_first, tile_step, s_count, forced_chunk, int(open_scene), resident_warps, max_strips,
                                      None if o is None else o.ctypes.data, counts.ctypes.data, plan.ctypes.data)
    keys = ("n_groups", "n_strips", "order_shift", "s_chunk", "s_chunk_tail", "s_tail_begin", "n_units")
    return counts, bad, dict(zip(keys, (int(v) for v in plan)))
