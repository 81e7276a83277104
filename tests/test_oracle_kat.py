"""Known-answer tests that pin the CPU oracle from first principles.

The reference (cbiffle/rtiow-rust) has no golden vectors, KATs or fixtures for its render path
(only two Vec3-indexing doctests, src/vec3.rs:221-228,270-277), and no Rust toolchain exists here
to run it, so the oracle is "parity unpinned"; these tests check each restated function against
values derivable by hand from the reference source cited next to each test.
"""
import math

import numpy as np
import pytest

F32 = np.float32
MAX, MIN = float(np.finfo(np.float32).max), float(np.finfo(np.float32).min)


def ulp_diff(a, b):
    a = np.asarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.asarray(b, np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7FFFFFFF), a)
    b = np.where(b < 0, -(b & 0x7FFFFFFF), b)
    return np.abs(a - b)


# ------------------------------------------------------------------ RNG
def test_philox_known_answers(oracle):
    # Random123 kat_vectors, philox4x32-10
    assert list(oracle.philox([0, 0], [0, 0, 0, 0])) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert list(oracle.philox([0xFFFFFFFF] * 2, [0xFFFFFFFF] * 4)) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert list(oracle.philox([0xA4093822, 0x299F31D0], [0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344])) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_u32_to_float(oracle):
    L = oracle.lib()
    assert L.oracle_u32_to_unit_f32(0) == 0.0
    assert L.oracle_u32_to_unit_f32(0xFFFFFFFF) == float(F32(1.0) - F32(2.0) ** -24)  # never 1.0
    assert L.oracle_u32_to_unit_f32(0x80000000) == 0.5
    assert L.oracle_u32_to_unit_f32(0x000000FF) == 0.0  # low 8 bits are discarded
    assert L.oracle_u32_to_f32_1_2(0) == 1.0
    assert L.oracle_u32_to_f32_1_2(0xFFFFFFFF) == float(F32(2.0) - F32(2.0) ** -23)


def test_smallrng_is_deterministic(oracle):
    a = oracle.smallrng_u32(0xDEADBEEF, 16)
    assert np.array_equal(a, oracle.smallrng_u32(0xDEADBEEF, 16))
    assert not np.array_equal(a, oracle.smallrng_u32(0xDEADBEF0, 16))
    assert len(set(a.tolist())) == 16


# ------------------------------------------------------------------ shared transcendentals
def test_log_f32(oracle):
    L = oracle.lib()
    assert L.oracle_log_f32(1.0) == 0.0
    assert L.oracle_log_f32(0.0) == -math.inf
    assert math.isnan(L.oracle_log_f32(-1.0))
    xs = np.concatenate([np.arange(1, 1 << 24, 4099, dtype=np.float64) / (1 << 24),  # every rng() value shape
                         np.float64(2.0) ** np.arange(-24, 10), np.linspace(0.5, 4, 2001)]).astype(np.float32)
    got = np.array([L.oracle_log_f32(float(x)) for x in xs], np.float32)
    want = np.log(xs.astype(np.float64)).astype(np.float32)  # correctly rounded via f64
    assert ulp_diff(got, want).max() <= 1
    assert (ulp_diff(got, want) == 0).mean() > 0.999


def test_sin_f32(oracle):
    L = oracle.lib()
    assert L.oracle_sin_f32(0.0) == 0.0
    xs = np.concatenate([np.linspace(-20000, 20000, 40001), np.linspace(-7, 7, 4001)]).astype(np.float32)
    got = np.array([L.oracle_sin_f32(float(x)) for x in xs], np.float32)
    want = np.sin(xs.astype(np.float64)).astype(np.float32)
    assert ulp_diff(got, want).max() <= 1
    assert math.isnan(L.oracle_sin_f32(math.inf))


def test_pow5_and_schlick(oracle):
    L = oracle.lib()
    for x in [0.0, 1.0, 0.5, 0.25, -0.5, 0.9999999, 1e-9]:
        assert L.oracle_pow5_f32(x) == float(F32(float(F32(x)) ** 5))
    # material.rs:142-146: r0 = ((1-1.5)/(1+1.5))^2 = 0.04 ; cos = 1 -> r0 exactly
    r0 = F32((F32(1) - F32(1.5)) / (F32(1) + F32(1.5)))
    r0 = F32(r0 * r0)
    assert L.oracle_schlick(1.0, 1.5) == float(r0)
    assert abs(L.oracle_schlick(1.0, 1.5) - 0.04) < 1e-7
    assert L.oracle_schlick(0.0, 1.5) == 1.0  # grazing: r0 + (1-r0)*1


# ------------------------------------------------------------------ vec3.rs
def test_vec3_order_of_operations(oracle):
    L = oracle.lib()
    a, b = [1e8, 1.0, -1e8], [1.0, 1.0, 1.0]
    # dot = (a0*b0 + a1*b1) + a2*b2 (vec3.rs:43-46,100-102): (1e8 + 1) rounds to 1e8 in f32 -> 0
    assert L.oracle_dot(oracle._f(a), oracle._f(b)) == 0.0
    assert L.oracle_dot(oracle._f([1.0, -1e8, 1e8]), oracle._f(b)) == 0.0  # (1 - 1e8) + 1e8
    assert L.oracle_dot(oracle._f([-1e8, 1e8, 1.0]), oracle._f(b)) == 1.0
    o = (oracle.C.c_float * 3)()
    L.oracle_cross(oracle._f([1, 0, 0]), oracle._f([0, 1, 0]), o)
    assert list(o) == [0.0, -0.0, 1.0]
    L.oracle_into_unit(oracle._f([3, 0, 4]), o)  # three divisions by length (vec3.rs:66-68,145-152)
    assert list(o) == [float(F32(3) / F32(5)), 0.0, float(F32(4) / F32(5))]


def test_reflect_refract(oracle):
    assert list(oracle.reflect([1, -1, 0], [0, 1, 0])) == [1.0, 1.0, 0.0]  # vec3.rs:313-315
    # normal incidence refracts straight through (vec3.rs:321-330)
    r = oracle.refract([0, -2, 0], [0, 1, 0], 1 / 1.5)
    assert np.allclose(r, [0, -1, 0])
    # total internal reflection: exiting glass (ni/nt = 1.5) beyond asin(1/1.5) = 41.81 deg
    def ray(deg):
        return [math.sin(math.radians(deg)), -math.cos(math.radians(deg)), 0]
    assert oracle.refract(ray(41.0), [0, 1, 0], 1.5) is not None
    assert oracle.refract(ray(42.5), [0, 1, 0], 1.5) is None
    # Snell: sin(out) = 1.5 sin(in)
    out = oracle.refract(ray(30.0), [0, 1, 0], 1.5)
    assert abs(out[0] / np.linalg.norm(out) - 0.75) < 1e-6


# ------------------------------------------------------------------ aabb.rs
def test_aabb_slab(oracle):
    mn, mx = [-1, -1, -1], [1, 1, 1]
    assert oracle.aabb_hit(mn, mx, [0, 0, -5, 0, 0, 1, 0], 0.001, MAX)
    assert not oracle.aabb_hit(mn, mx, [0, 0, -5, 0, 0, -1, 0], 0.001, MAX)       # behind
    assert oracle.aabb_hit(mn, mx, [0, 0, 5, 0, 0, -1, 0], 0.001, MAX)            # negative inv_d swaps
    assert not oracle.aabb_hit(mn, mx, [0, 0, -5, 0, 0, 1, 0], 0.001, 4.0)        # t range ends at entry: end > start fails
    assert oracle.aabb_hit(mn, mx, [0, 0, -5, 0, 0, 1, 0], 0.001, 4.0001)
    assert not oracle.aabb_hit(mn, mx, [0, 0, -5, 0, 0, 1, 0], 6.0, MAX)          # starts at exit
    assert not oracle.aabb_hit(mn, mx, [2, 0, -5, 0, 0, 1, 0], 0.001, MAX)        # zero dir component, outside slab
    assert oracle.aabb_hit(mn, mx, [0.5, 0, -5, 0, 0, 1, 0], 0.001, MAX)          # zero dir component, inside slab
    # origin exactly on a slab plane with zero direction: 0 * inf = NaN, ignored by f32::max/min
    assert oracle.aabb_hit(mn, mx, [1, 0, -5, 0, 0, 1, 0], 0.001, MAX)
    assert oracle.aabb_hit(mn, mx, [1, 0, -5, -0.0, 0, 1, 0], 0.001, MAX)
    assert not oracle.aabb_hit(mn, mx, [-3, -3, -3, 1, 1, -1, 0], 0.001, MAX)


# ------------------------------------------------------------------ object.rs primitives
def test_sphere_roots_and_range_edges(oracle):
    ray = [0, 0, -5, 0, 0, 1, 0]
    h = oracle.sphere_hit(1.0, ray, 0.001, MAX)
    assert h[0] == 4.0 and list(h[1:4]) == [0, 0, -1] and list(h[4:7]) == [0, 0, -1]   # object.rs:96,104
    assert oracle.sphere_hit(1.0, ray, 0.001, 4.0) is None        # t < end is strict; far root 6 also fails
    assert oracle.sphere_hit(1.0, ray, 4.0, MAX)[0] == 4.0        # t >= start is inclusive
    assert oracle.sphere_hit(1.0, ray, 4.5, MAX)[0] == 6.0        # falls through to the far root (object.rs:97)
    assert oracle.sphere_hit(1.0, ray, 4.5, 6.0) is None
    assert oracle.sphere_hit(1.0, [0, 1, -5, 0, 0, 1, 0], 0.001, MAX) is None  # tangent: discriminant == 0 is a miss
    # un-normalised direction: t scales inversely, normal = p / radius in the sphere's frame
    h = oracle.sphere_hit(2.0, [0, 0, -5, 0, 0, 2, 0], 0.001, MAX)
    assert h[0] == 1.5 and list(h[4:7]) == [0, 0, -1]
    # from inside: near root negative, far root accepted
    assert oracle.sphere_hit(1.0, [0, 0, 0, 0, 0, 1, 0], 0.001, MAX)[0] == 1.0


def test_rect_half_open_edges(oracle):
    # Rect orthogonal to Z: OTHER1 = X, OTHER2 = Y (object.rs:177-181); ranges are half-open (:201-207)
    def hit(x, y, t0=0.001, t1=MAX, flip=False):
        return oracle.rect_hit(2, (0.0, 1.0), (0.0, 2.0), 3.0, [x, y, 0, 0, 0, 1, 0], t0, t1, flip)
    h = hit(0.5, 0.5)
    assert h[0] == 3.0 and list(h[4:7]) == [0, 0, 1]
    assert hit(0.0, 0.0) is not None and hit(1.0, 0.5) is None and hit(0.5, 2.0) is None
    assert hit(-1e-7, 0.5) is None
    assert hit(0.5, 0.5, 3.0, MAX) is not None and hit(0.5, 0.5, 0.001, 3.0) is None  # t in [start, end)
    assert list(hit(0.5, 0.5, flip=True)[4:7]) == [-0.0, -0.0, -1.0]  # FlipNormals (object.rs:249-252)
    # normal is +axis regardless of the side the ray comes from (object.rs:210-211)
    h = oracle.rect_hit(1, (0, 1), (0, 1), 0.0, [0.5, 5, 0.5, 0, -1, 0, 0], 0.001, MAX)
    assert list(h[4:7]) == [0, 1, 0]
    # X-orthogonal: OTHER1 = Y, OTHER2 = Z
    assert oracle.rect_hit(0, (0, 1), (10, 11), 2.0, [0, 0.5, 10.5, 1, 0, 0, 0], 0.001, MAX) is not None
    assert oracle.rect_hit(0, (0, 1), (10, 11), 2.0, [0, 10.5, 0.5, 1, 0, 0, 0], 0.001, MAX) is None
    # parallel ray: t = k/0 = inf (or NaN) -> rejected
    assert oracle.rect_hit(2, (0, 1), (0, 1), 3.0, [0.5, 0.5, 0, 1, 0, 0, 0], 0.001, MAX) is None


def test_wrappers(oracle):
    ray = [0, 0, -5, 0, 0, 1, 0.5]
    h = oracle.wrapped_sphere_hit(oracle.WRAP_TRANSLATE, [0, 0, 2], 1.0, ray, 0.001, MAX)
    assert h[0] == 6.0 and list(h[1:4]) == [0, 0, 1] and list(h[4:7]) == [0, 0, -1]      # object.rs:275-282
    h = oracle.wrapped_sphere_hit(oracle.WRAP_FLIP, [0, 0, 0], 1.0, ray, 0.001, MAX)
    assert list(h[4:7]) == [-0.0, -0.0, 1.0]
    # LinearMove quirk: p stays in the moving frame (object.rs:498-512 has no .map on the result)
    h = oracle.wrapped_sphere_hit(oracle.WRAP_LINEAR_MOVE, [0, 0, 2], 1.0, ray, 0.001, MAX)
    assert h[0] == 5.0 and list(h[1:4]) == [0, 0, -1]   # world point would be (0,0,0)
    # Scale (object.rs:301-319): o/f, d/f in; p*f, n/f out
    h = oracle.wrapped_sphere_hit(oracle.WRAP_SCALE, [1, 1, 2], 1.0, ray, 0.001, MAX)
    assert h[0] == 3.0 and list(h[1:4]) == [0, 0, -2] and list(h[4:7]) == [0, 0, -0.5]
    # RotateY by 90 degrees is a no-op for a centred sphere's t, and rotates p back (object.rs:341-370)
    h = oracle.wrapped_sphere_hit(oracle.WRAP_ROTATE_Y, [90, 0, 0], 1.0, ray, 0.001, MAX)
    assert abs(h[0] - 4.0) < 1e-5 and np.allclose(h[1:4], [0, 0, -1], atol=1e-6)


def test_rotate_y_direction_convention(oracle):
    # rot(p,s,c) = (p.(c,0,s), p.y, p.(-s,0,c)) (object.rs:349-355).  A sphere translated to +x inside a
    # RotateY(90) ends up at world -z ... checked via the kitchen-sink free function on a Scene:
    # use the oracle's hit_top on `cornell`: the tall box is rotated +15 degrees about Y.
    sc = oracle.Scene("cornell", 64, 64, top_level_bvh=False)
    # straight down onto the tall box's top face (y = 330) near its centre
    c, s = math.cos(math.radians(15)), math.sin(math.radians(15))
    lx, lz = 82.5, 82.5
    wx, wz = c * lx + s * lz + 265, -s * lx + c * lz + 295
    h = sc.hit_top([wx, 500, wz, 0, -1, 0, 0])
    assert h is not None and abs(h[2] - 330) < 1e-3 and np.allclose(h[4:7], [0, 1, 0], atol=1e-6)
    # just outside the rotated footprint the ray reaches the floor instead
    lx = -2.0
    wx, wz = c * lx + s * lz + 265, -s * lx + c * lz + 295
    h = sc.hit_top([wx, 500, wz, 0, -1, 0, 0])
    assert abs(h[2]) < 1e-3


def test_constant_medium(oracle):
    ray = [0, 0, -5, 0, 0, 2, 0]  # |d| = 2; boundary sphere r=1: t in (2, 3)
    # u -> 1: hit_distance -> 0+ : scatters right at the entry point (object.rs:562-565)
    u1 = float(F32(1) - F32(2) ** -24)
    t = oracle.medium_hit_with_u(1.0, 1.0, ray, 0.001, MAX, u1)
    assert t is not None and abs(t - 2.0) < 1e-6
    # u = 0: ln(0) = -inf -> distance +inf -> never
    assert oracle.medium_hit_with_u(1.0, 1e9, ray, 0.001, MAX, 0.0) is None
    # distance_inside = (3-2)*2 = 2 ; density 1 -> hits iff -ln(u) < 2 ; t = 2 + (-ln u)/2
    u = math.exp(-1.0)
    t = oracle.medium_hit_with_u(1.0, 1.0, ray, 0.001, MAX, u)
    assert abs(t - 2.5) < 1e-6
    assert oracle.medium_hit_with_u(1.0, 1.0, ray, 0.001, MAX, math.exp(-2.1)) is None
    # t_range clamps the interval (object.rs:553-557)
    assert oracle.medium_hit_with_u(1.0, 1.0, ray, 0.001, 2.0, u) is None      # hit1.t >= hit2.t
    assert oracle.medium_hit_with_u(1.0, 1.0, ray, 0.001, 2.25, u) is None     # inside length 0.5 < 1
    # ray starting inside: hit1 is behind the origin, clamped to t_start
    t = oracle.medium_hit_with_u(1.0, 1.0, [0, 0, 0, 0, 0, 1, 0], 0.001, MAX, math.exp(-0.5))
    assert abs(t - 0.501) < 1e-6


# ------------------------------------------------------------------ perlin.rs / texture.rs
def test_perlin(oracle):
    sc = oracle.Scene("final", 8, 8)
    vecs, perms = sc.perlin()
    assert (np.linalg.norm(vecs, axis=1) < 1).all()          # un-normalised in_unit_sphere vectors (perlin.rs:15-21)
    for p in perms:
        assert sorted(p.tolist()) == list(range(256))         # permutations (perlin.rs:5-13)
    for p in [(0, 0, 0), (1, 2, 3), (-4, 7, 255), (256, -256, 13)]:
        assert sc.perlin_noise(p) == 0.0                     # weight vectors vanish on lattice points
    assert sc.perlin_noise((0.5, 0.25, 0.75)) == sc.perlin_noise((256.5, 0.25, -255.25))  # period 256
    assert abs(sc.perlin_noise((0.3, 0.6, 0.9))) < 1.0
    t = sc.perlin_turb((0.3, 0.6, 0.9))
    assert t >= 0 and t == sc.perlin_turb((0.3, 0.6, 0.9))
    # turb = |sum_k 2^-k noise(2^k p)| (perlin.rs:66-75)
    acc, w, p = F32(0), F32(1), np.array([0.3, 0.6, 0.9], np.float32)
    for _ in range(7):
        acc = F32(acc + F32(w * F32(sc.perlin_noise(p))))
        w = F32(w * F32(0.5))
        p = (F32(2) * p).astype(np.float32)
    assert t == float(abs(acc))


def test_checker(oracle):
    c0, c1 = [1, 0, 0], [0, 0, 1]
    # s = sin(10x) sin(10y) sin(10z) ; s < 0 -> t1 else t0 (texture.rs:12-21)
    assert list(oracle.checker([0.1, 0.1, 0.1], c0, c1)) == c0       # all sines positive
    assert list(oracle.checker([0.4, 0.1, 0.1], c0, c1)) == c1       # sin(4) < 0
    assert list(oracle.checker([0.4, 0.4, 0.1], c0, c1)) == c0
    assert list(oracle.checker([0.0, 0.1, 0.1], c0, c1)) == c0       # s == 0 -> t0


# ------------------------------------------------------------------ camera.rs
def test_camera_look(oracle):
    cam = oracle.camera_look([13, 2, 3], [0, 0, 0], [0, 1, 0], 20.0, 2.0, 0.1, 10.0)
    origin, llc, hor, ver, u, v = (cam[3 * i:3 * i + 3] for i in range(6))
    assert list(origin) == [13, 2, 3] and cam[18] == F32(0.05) and cam[19] == 0 and cam[20] == 1
    w = np.array([13, 2, 3]) / math.sqrt(182)
    assert np.allclose(np.cross(u, v), w, atol=1e-6) and abs(np.dot(u, v)) < 1e-6
    hh = math.tan(math.radians(10.0))
    assert abs(np.linalg.norm(ver) - 2 * hh * 10) < 1e-4 and abs(np.linalg.norm(hor) - 2 * 2 * hh * 10) < 1e-4
    centre = llc + 0.5 * hor + 0.5 * ver
    assert np.allclose(centre - origin, -10 * w, atol=1e-4)       # focus plane centre is 10 units toward look_at


def test_get_ray(oracle):
    sc = oracle.Scene("cornell", 40, 40)    # aperture 0: lens offset is exactly zero but still drawn
    cam = sc.camera()
    r = sc.get_ray(3, 7, 2)
    assert list(r[0:3]) == [278, 278, -800] and 0 <= r[6] < 1
    assert not np.array_equal(r, sc.get_ray(3, 7, 3)) and not np.array_equal(r, sc.get_ray(4, 7, 2))
    assert np.array_equal(r, sc.get_ray(3, 7, 2))
    sb = oracle.Scene("book1", 40, 20)
    rb = sb.get_ray(0, 0, 0)
    off = rb[0:3] - np.array([13, 2, 3], np.float32)
    assert 0 < np.linalg.norm(off) < 0.05 + 1e-6               # inside the lens disc of radius aperture/2
    # origin + direction lands on the focus plane point regardless of the lens offset (camera.rs:58-60)
    cb = sb.camera()
    on_plane = rb[0:3] + rb[3:6] - cb[3:6]
    a = np.linalg.lstsq(np.stack([cb[6:9], cb[9:12]], 1).astype(np.float64), on_plane.astype(np.float64), rcond=None)[0]
    assert 0 <= a[0] < 1 / 40 + 1e-5 and 0 <= a[1] < 1 / 20 + 1e-5     # pixel (0,0) jittered inside its cell


# ------------------------------------------------------------------ lib.rs: cast / color / print_ppm
def test_to_u8(oracle):
    L = oracle.lib()
    assert [L.oracle_to_u8(x) for x in (-1.0, 0.0, 0.5, 1.0, 2.0, 1e30, -1e30)] == [0, 0, 127, 255, 255, 255, 0]
    assert L.oracle_to_u8(math.nan) == 0 and L.oracle_to_u8(math.inf) == 255
    q = oracle.ppm_quantise(np.array([0.25, 1.0, 0.0], np.float32))      # sqrt gamma first (lib.rs:348)
    assert q.tolist() == [127, 255, 0]


def test_image_orientation_and_sky(oracle):
    sc = oracle.Scene("book1", 32, 16)
    img, _, _ = sc.render(2)
    assert img.shape == (16, 32, 3)
    top, bottom = img[0].mean(0), img[-1].mean(0)
    assert top[2] > top[0] and top[2] > 0.8          # row 0 is the top scanline: blue-white sky (lib.rs:326-330)
    assert abs(bottom[0] - bottom[2]) < 0.25         # bottom: grey ground
    black, _, _ = sc.render(2, background=0)
    assert black[0].max() == 0.0                      # HEAD semantics: escaped rays are black (lib.rs:100)


def test_cast_equals_par_cast_and_row_ranges(oracle):
    sc = oracle.Scene("cornell", 24, 24, top_level_bvh=False)
    seq, _, _ = sc.render(4, nthreads=1)
    par, _, _ = sc.render(4, nthreads=4)
    assert np.array_equal(seq, par)
    part, _, _ = sc.render(4, rows=(5, 11), nthreads=2)
    assert np.array_equal(part, seq[5:11])
    other, _, _ = sc.render(4, seed=1)
    assert not np.array_equal(other, seq)


def test_list_and_bvh_top_level_agree(oracle):
    # USE_BVH (main.rs:321) must not change the picture: nearest-hit is independent of the container
    for name in ("cornell", "final", "kitchen_sink"):
        a, _, _ = oracle.Scene(name, 20, 20, top_level_bvh=False).render(4, nthreads=4)
        b, _, _ = oracle.Scene(name, 20, 20, top_level_bvh=True).render(4, nthreads=4)
        assert np.array_equal(a, b), name


def test_sample_sum_is_left_fold(oracle):
    sc = oracle.Scene("book1", 16, 8)
    img, smp, _ = sc.render(5, want_samples=True)
    acc = np.zeros_like(img)
    for s in range(5):                                  # iter.fold(Vec3::default(), Add) (vec3.rs:195-203)
        acc = (acc + smp[:, :, s, :]).astype(np.float32)
    assert np.array_equal(img, (acc / F32(5)).astype(np.float32))      # col / ns as f32 (lib.rs:374)


def test_bounce_cap_and_counters(oracle):
    sc = oracle.Scene("cornell", 48, 48, top_level_bvh=False)
    _, _, c = sc.render(16, nthreads=4, want_counters=True)
    assert c["samples"] == 48 * 48 * 16
    assert c["max_segments"] == 51                      # bounces 0..=50 -> 51 hit_top calls (lib.rs:93-97)
    assert c["rect_tests"] == 18 * c["segments"]        # 6 walls + 2 prisms of 6: every object, every segment (lib.rs:40)
    assert c["node_tests"] == 0 and c["sphere_tests"] == 0
    b = oracle.algorithmic_bytes_per_sample(c, 16)
    assert abs(b - (32 * c["rect_tests"] / c["samples"] + 12 / 16)) < 1e-9


def test_bvh_shape(oracle):
    sc = oracle.Scene("book1", 8, 8)
    info = sc.info()
    n = info["n_top_objects"]
    assert 470 <= n <= 488 and info["bvh_nodes"] == 2 * n - 1      # one object per leaf (bvh.rs:61-65)
    assert info["bvh_depth"] == math.ceil(math.log2(n)) + 1        # median split (bvh.rs:68-72)
    with pytest.raises(RuntimeError):
        oracle.Scene("nope", 8, 8)
