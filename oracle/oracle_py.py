"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding over oracle/_build/liboracle.so (the CPU restatement of cbiffle/rtiow-rust's
render path).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module; nothing under rtiow-rust_b200/ does.

PARITY UNPINNED: the reference has no golden vectors for this path and cannot be built here
(no Rust toolchain), see DESIGN.md.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

COUNTER_NAMES = ("samples", "segments", "node_tests", "sphere_tests", "rect_tests", "medium_evals", "draws",
                 "max_segments")


def build(force=False):
    """Compile the oracle with the committed Makefile (g++ only; no GPU needed)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-j4"] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        fp = C.POINTER(C.c_float)
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_scene_build.restype = C.c_void_p
        L.oracle_scene_build.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int]
        L.oracle_scene_free.argtypes = [C.c_void_p]
        L.oracle_render.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int, C.c_uint32,
                                    C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_ppm_quantise.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.oracle_scene_info.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_scene_camera.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_scene_perlin.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        for name in ("oracle_u32_to_unit_f32", "oracle_u32_to_f32_1_2"):
            getattr(L, name).restype = C.c_float
            getattr(L, name).argtypes = [C.c_uint32]
        for name in ("oracle_log_f32", "oracle_sin_f32", "oracle_pow5_f32"):
            getattr(L, name).restype = C.c_float
            getattr(L, name).argtypes = [C.c_float]
        L.oracle_schlick.restype = C.c_float
        L.oracle_schlick.argtypes = [C.c_float, C.c_float]
        L.oracle_to_u8.argtypes = [C.c_float]
        L.oracle_dot.restype = C.c_float
        L.oracle_dot.argtypes = [fp, fp]
        L.oracle_cross.argtypes = [fp, fp, fp]
        L.oracle_into_unit.argtypes = [fp, fp]
        L.oracle_reflect.argtypes = [fp, fp, fp]
        L.oracle_refract.argtypes = [fp, fp, C.c_float, fp]
        L.oracle_aabb_hit.argtypes = [fp, fp, fp, C.c_float, C.c_float]
        L.oracle_sphere_hit.argtypes = [C.c_float, fp, C.c_float, C.c_float, fp]
        L.oracle_rect_hit.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, fp,
                                      C.c_float, C.c_float, fp]
        L.oracle_wrapped_sphere_hit.argtypes = [C.c_int, fp, C.c_float, fp, C.c_float, C.c_float, fp]
        L.oracle_medium_hit_with_u.argtypes = [C.c_float, C.c_float, fp, C.c_float, C.c_float, C.c_float, fp]
        L.oracle_perlin_noise.restype = C.c_float
        L.oracle_perlin_noise.argtypes = [C.c_void_p, fp]
        L.oracle_perlin_turb.restype = C.c_float
        L.oracle_perlin_turb.argtypes = [C.c_void_p, fp, C.c_int]
        L.oracle_checker.argtypes = [fp, fp, fp, fp]
        L.oracle_camera_look.argtypes = [fp, fp, fp] + [C.c_float] * 6 + [fp]
        L.oracle_get_ray.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                     C.c_uint64, fp]
        L.oracle_hit_top.argtypes = [C.c_void_p, fp, fp]
        L.oracle_smallrng_u32.argtypes = [C.c_uint64, C.c_uint32, C.c_void_p]
        _lib = L
    return _lib


def _f(vals):
    return (C.c_float * len(vals))(*[float(v) for v in vals])


F32_MAX = float(np.finfo(np.float32).max)
F32_MIN = float(np.finfo(np.float32).min)


class Scene:
    """A builtin scene by name (see oracle_scenes.hpp: scenes::build)."""

    def __init__(self, name, nx, ny, scene_seed=0xDEADBEEF, top_level_bvh=True):
        self.name, self.nx, self.ny = name, nx, ny
        self._h = lib().oracle_scene_build(name.encode(), nx, ny, scene_seed, int(top_level_bvh))
        if not self._h:
            raise RuntimeError(lib().oracle_last_error().decode())

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.oracle_scene_free(self._h)
            self._h = None

    def render(self, ns, seed=0xDEADBEEF, background=-1, rows=None, nthreads=1, want_samples=False,
               want_counters=False):
        """Returns (image[rows, nx, 3] float32 with row 0 = top, samples or None, counters dict or None)."""
        r0, r1 = rows if rows is not None else (0, self.ny)
        img = np.zeros((r1 - r0, self.nx, 3), np.float32)
        smp = np.zeros((r1 - r0, self.nx, ns, 3), np.float32) if want_samples else None
        cnt = np.zeros(8, np.uint64) if want_counters else None
        rc = lib().oracle_render(self._h, self.nx, self.ny, ns, seed, background, r0, r1, nthreads,
                                 img.ctypes.data, smp.ctypes.data if want_samples else None,
                                 cnt.ctypes.data if want_counters else None)
        if rc:
            raise RuntimeError(lib().oracle_last_error().decode())
        counters = dict(zip(COUNTER_NAMES, (int(v) for v in cnt))) if want_counters else None
        return img, smp, counters

    def info(self):
        out = np.zeros(5, np.uint64)
        lib().oracle_scene_info(self._h, out.ctypes.data)
        return dict(top_is_bvh=bool(out[0]), n_top_objects=int(out[1]), bvh_nodes=int(out[2]), bvh_depth=int(out[3]),
                    n_media=int(out[4]))

    def camera(self):
        out = np.zeros(21, np.float32)
        lib().oracle_scene_camera(self._h, out.ctypes.data)
        return out

    def perlin(self):
        vecs = np.zeros((256, 3), np.float32)
        perms = np.zeros((3, 256), np.uint8)
        lib().oracle_scene_perlin(self._h, vecs.ctypes.data, perms.ctypes.data)
        return vecs, perms

    def get_ray(self, x, y, s, seed=0xDEADBEEF):
        out = (C.c_float * 7)()
        if lib().oracle_get_ray(self._h, self.nx, self.ny, x, y, s, seed, out):
            raise RuntimeError(lib().oracle_last_error().decode())
        return np.array(out[:], np.float32)

    def hit_top(self, ray):
        out = (C.c_float * 7)()
        ok = lib().oracle_hit_top(self._h, _f(ray), out)
        return np.array(out[:], np.float32) if ok else None

    def perlin_noise(self, p):
        return lib().oracle_perlin_noise(self._h, _f(p))

    def perlin_turb(self, p, depth=7):
        return lib().oracle_perlin_turb(self._h, _f(p), depth)


def ppm_quantise(linear):
    """print_ppm's per-channel sqrt + to_u8 (src/lib.rs:344-361) -> int32 array of the same shape."""
    a = np.ascontiguousarray(linear, np.float32)
    out = np.zeros(a.shape, np.int32)
    lib().oracle_ppm_quantise(a.ctypes.data, a.size, out.ctypes.data)
    return out


def algorithmic_bytes_per_sample(counters, ns):
    """B = 32*N_node + 32*N_sphere + 32*N_rect + 16*N_medium (+12/ns) per sample,
    counted on the reference's own left-first traversal."""
    n = counters["samples"]
    return (32.0 * counters["node_tests"] + 32.0 * counters["sphere_tests"] + 32.0 * counters["rect_tests"] +
            16.0 * counters["medium_evals"]) / n + 12.0 / ns


def philox(key, ctr):
    k = np.array(key, np.uint32)
    c = np.array(ctr, np.uint32)
    o = np.zeros(4, np.uint32)
    lib().oracle_philox4x32_10(k.ctypes.data, c.ctypes.data, o.ctypes.data)
    return o


def sphere_hit(radius, ray, t0, t1):
    out = (C.c_float * 7)()
    ok = lib().oracle_sphere_hit(radius, _f(ray), t0, t1, out)
    return np.array(out[:], np.float32) if ok else None


def rect_hit(axis, r0, r1, k, ray, t0, t1, flip=False):
    out = (C.c_float * 7)()
    ok = lib().oracle_rect_hit(axis, r0[0], r0[1], r1[0], r1[1], k, int(flip), _f(ray), t0, t1, out)
    return np.array(out[:], np.float32) if ok else None


WRAP_TRANSLATE, WRAP_SCALE, WRAP_ROTATE_Y, WRAP_LINEAR_MOVE, WRAP_FLIP = range(5)


def wrapped_sphere_hit(wrapper, v, radius, ray, t0, t1):
    out = (C.c_float * 7)()
    ok = lib().oracle_wrapped_sphere_hit(wrapper, _f(v), radius, _f(ray), t0, t1, out)
    return np.array(out[:], np.float32) if ok else None


def aabb_hit(mn, mx, ray, t0, t1):
    return bool(lib().oracle_aabb_hit(_f(mn), _f(mx), _f(ray), t0, t1))


def medium_hit_with_u(radius, density, ray, t0, t1, u):
    t = C.c_float()
    ok = lib().oracle_medium_hit_with_u(radius, density, _f(ray), t0, t1, u, C.byref(t))
    return t.value if ok else None


def reflect(v, n):
    o = (C.c_float * 3)()
    lib().oracle_reflect(_f(v), _f(n), o)
    return np.array(o[:], np.float32)


def refract(v, n, ni_over_nt):
    o = (C.c_float * 3)()
    ok = lib().oracle_refract(_f(v), _f(n), ni_over_nt, o)
    return np.array(o[:], np.float32) if ok else None


def checker(p, c0, c1):
    o = (C.c_float * 3)()
    lib().oracle_checker(_f(p), _f(c0), _f(c1), o)
    return np.array(o[:], np.float32)


def camera_look(look_from, look_at, up, fov, aspect, aperture, focus, e0=0.0, e1=1.0):
    o = (C.c_float * 21)()
    lib().oracle_camera_look(_f(look_from), _f(look_at), _f(up), fov, aspect, aperture, focus, e0, e1, o)
    return np.array(o[:], np.float32)


def smallrng_u32(seed, n):
    o = np.zeros(n, np.uint32)
    lib().oracle_smallrng_u32(seed, n, o.ctypes.data)
    return o
