"""CPU-side tests of the product's host layer (no GPU): the C ABI library loads and exports every
declared symbol, scene validation, the C++ host mirror's flattener, and — through the host-compiled
copy of the kernel's per-path code (tests/kernel_host_harness.cpp) — that flattened scenes traverse
and shade exactly like the oracle's object tree."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import rtiow_rust_b200 as R
from rtiow_rust_b200 import _native as N

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness_lib as H  # noqa: E402
from common import bits_equal, golden_cases, n_diff  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "rtiow_b200.h")).read()
    declared = set(re.findall(r"\b(rtiow_b200_\w+)\s*\(", header))
    assert declared == set(N.ABI_SYMBOLS)
    for flavour, path in (("parity", N.ABI_LIB), ("fast", N.FAST_ABI_LIB)):   # the default build and the tolerance build (make FAST=1)
        lib = N.abi(flavour)
        for sym in declared:
            assert getattr(lib, sym) is not None
        assert lib.rtiow_b200_abi_version() == 2
        assert lib.rtiow_b200_build_flavour().decode().startswith(flavour)
        nm = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True).stdout
        for sym in declared:
            assert re.search(rf"\bT {sym}\b", nm), sym
        exported = set(re.findall(r"\bT (\w+)", nm))
        assert exported == declared, exported ^ declared       # nothing else leaks out of the library


def test_header_is_plain_c_and_links(tmp_path):
    """The header as a C11 translation unit (-Wall -Wextra -Werror -pedantic) linked against the shipped library: what a
    cgo / bindgen / JNI binding would read.  Runs the GPU-free entry points on a hand-written one-sphere descriptor."""
    exe = str(tmp_path / "abi_smoke")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-o", exe,
                           os.path.join(ROOT, "tests", "c_abi", "abi_smoke.c"), "-L" + N.BUILD_DIR, "-lrtiow_b200",
                           "-Wl,-rpath," + N.BUILD_DIR])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "abi_smoke ok" in out.stdout, (out.returncode, out.stdout, out.stderr)


def test_abi_struct_sizes_match_header():
    assert C.sizeof(N.Item) == 32 and C.sizeof(N.XformOp) == 16 and C.sizeof(N.Frame) == 8
    assert C.sizeof(N.MaterialRec) == 32 and C.sizeof(N.TextureRec) == 32 and C.sizeof(N.CameraRec) == 84


def test_kernels_are_compiled_for_sm_100a_with_tma():
    log = open(os.path.join(N.BUILD_DIR, "ptxas.log")).read()
    assert "for 'sm_100a'" in log and "render_kernel" in log
    # the 256-thread execution shapes have registers to spare and must not spill; the 512-thread ones (128 registers) may
    # spill a few words of the general kernel's cold paths, the register-capped 768/1024-thread variants more.  Template: <kSmem, kFrames, kFast, kFeat, kThreads, kMinBlocks>
    checked = 0
    for m in re.finditer(r"Function properties for (\S+)\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", log):
        k = re.search(r"render_kernelILb[01]ELb[01]ELb[01]ELj\d+ELi(\d+)ELi\d+E", m.group(1))
        if not k:
            continue
        limit = {256: 0, 512: 256}.get(int(k.group(1)), 768)
        assert int(m.group(3)) <= limit, (m.group(1), m.group(3))   # a runaway spill is a perf bug
        checked += int(k.group(1)) == 256
    assert checked >= 8, "the spill check matched no kernel: the mangled-name pattern is stale"
    sass = subprocess.run(["cuobjdump", "-sass", N.ABI_LIB], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass          # cp.async.bulk (TMA bulk copy) staging of the scene blob
    assert "SYNCS" in sass           # mbarrier expect_tx / try_wait
    assert "STG.E.128" in sass       # 16-byte staging store per pixel-sample
    assert "VOTE" in sass            # __ballot_sync job hand-out


def test_no_gpu_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    world, cam = R.build_scene("cornell", 8, 8, use_bvh=False)
    with pytest.raises(R.RtiowError) as e:
        R.par_cast(8, 8, 1, cam, world)
    assert e.value.code == N.ERR_NO_DEVICE and "no CPU fallback" in str(e.value)


def test_product_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "rtiow-rust_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cpp", ".hpp", ".cu", ".cuh", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(#include|import|from)\b.*oracle", text, re.M), os.path.join(dirpath, f)
                assert "liboracle" not in text


@pytest.mark.parametrize("name", R.SCENES)
def test_every_scene_flattens_and_validates(name):
    for bvh in (False, True):
        world, cam = R.build_scene(name, 32, 32, use_bvh=bvh)
        world.validate()
        c = world.counts()
        items = world.items()
        assert (items[-1, 3] & 15) == 0                       # END
        kinds = items[:, 3] & 15
        skips = items[kinds == 1, 3] >> 4
        assert (skips > np.nonzero(kinds == 1)[0]).all()      # skip links only point forward
        if bvh:
            assert c["bbox"] >= 2 * len(world) - 1


def test_book1_stream_shape():
    world, _ = R.build_scene("book1", 400, 200)
    c = world.counts()
    n = len(world)
    assert c["spheres"] == n and c["bbox"] == 2 * n - 1 and c["items"] == 3 * n      # + END - 1
    assert c["frames"] == 1 and c["ops"] == 0      # every Translate{Sphere} is folded into its item
    assert c["background"] == 1


def test_final_scene_stream_shape():
    world, _ = R.build_scene("final", 64, 64, use_bvh=False)
    c = world.counts()
    assert c["rects"] == 2401 and c["spheres"] == 1007 and c["media"] == 2 and c["set_frames"] == 2
    assert c["bbox"] == 799 + 1999


def _desc_copy(world):
    d = N.SceneDesc()
    C.memmove(C.byref(d), world.desc, C.sizeof(d))
    return d


def test_validation_rejects_malformed_scenes():
    world, _ = R.build_scene("cornell", 16, 16, use_bvh=True)
    lib = N.abi()

    def expect(desc, code, fragment):
        assert lib.rtiow_b200_scene_validate(C.byref(desc)) == code
        assert fragment in lib.rtiow_b200_last_error().decode()

    d = _desc_copy(world)
    assert lib.rtiow_b200_scene_validate(C.byref(d)) == 0
    d.abi_version = 99
    expect(d, N.ERR_INVALID_ARG, "abi_version")
    d = _desc_copy(world)
    d.n_items -= 1                                      # drops END
    expect(d, N.ERR_INVALID_SCENE, "END")
    d = _desc_copy(world)
    d.n_items = 0
    expect(d, N.ERR_INVALID_SCENE, "zero items")        # cf. "Can't create a BVH from zero objects." (bvh.rs:60)
    # a backward skip link would let traversal loop forever on the device
    items = (N.Item * world.desc.contents.n_items)()
    C.memmove(items, world.desc.contents.items, C.sizeof(items))
    d = _desc_copy(world)
    d.items = items
    assert (items[0].a_w & 15) == 1
    items[0].a_w = 1 | (0 << 4)
    expect(d, N.ERR_INVALID_SCENE, "forward")
    C.memmove(items, world.desc.contents.items, C.sizeof(items))
    rect = next(i for i in range(len(items)) if (items[i].a_w & 15) == 3)
    items[rect].b_w = (items[rect].b_w & 0xFF000000) | 0x00FFFFFF
    expect(d, N.ERR_INVALID_SCENE, "material")
    assert lib.rtiow_b200_scene_validate(None) == N.ERR_INVALID_ARG


def test_unknown_scene_and_camera_look():
    with pytest.raises(RuntimeError):
        R.build_scene("nope", 8, 8)
    cam = R.Camera.look((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 2.0, 0.1, 10.0).as_array()
    assert list(cam[0:3]) == [13, 2, 3] and cam[18] == np.float32(0.05)


def test_camera_matches_oracle(oracle):
    for name in ("book1", "cornell", "final", "kitchen_sink"):
        _, cam = R.build_scene(name, 120, 80)
        assert np.array_equal(cam.as_array(), oracle.Scene(name, 120, 80).camera()), name
    a = R.Camera.look((1, 2, 3), (4, -5, 6), (0, 1, 0), 33.0, 1.5, 0.3, 7.0, (0.25, 0.75)).as_array()
    b = oracle.camera_look((1, 2, 3), (4, -5, 6), (0, 1, 0), 33.0, 1.5, 0.3, 7.0, 0.25, 0.75)
    assert np.array_equal(a, b)


def test_oracle_still_matches_golden(oracle):
    for key, name, nx, ny, ns, bvh, want, segs in golden_cases():
        got, _, cnt = oracle.Scene(name, nx, ny, top_level_bvh=bvh).render(ns, nthreads=4, want_counters=True)
        assert bits_equal(got, want), key
        assert cnt["segments"] == segs


def test_flattened_stream_logic_matches_golden():
    """The kernel's per-path code, compiled for the host, over the product's flattened scenes."""
    for key, name, nx, ny, ns, bvh, want, segs in golden_cases():
        world, cam = R.build_scene(name, nx, ny, use_bvh=bvh)
        for accel in (1, 2, 0):          # fast re-indexed traversal, exact re-indexed traversal, reference-order stream
            got, smp = H.render(world, cam, nx, ny, ns, want_samples=True, accel=accel)
            assert n_diff(got, want) == 0, (key, accel)
            assert int(smp[..., 3].sum()) == segs, (key, accel)


@pytest.mark.parametrize("name,bvh", [("book1", True), ("cornell", False), ("final", False), ("final", True),
                                      ("kitchen_sink", True), ("kitchen_sink", False), ("volume_test", False)])
@pytest.mark.parametrize("accel", [1, 2, 0])
def test_flattened_stream_logic_matches_oracle_per_sample(oracle, name, bvh, accel):
    nx, ny, ns = 40, 30, 5
    world, cam = R.build_scene(name, nx, ny, use_bvh=bvh)
    _, smp = H.render(world, cam, nx, ny, ns, seed=12345, want_samples=True, accel=accel)
    _, osmp, cnt = oracle.Scene(name, nx, ny, top_level_bvh=bvh).render(ns, seed=12345, nthreads=4, want_samples=True,
                                                                       want_counters=True)
    assert n_diff(smp[..., :3], osmp) == 0
    assert int(smp[..., 3].sum()) == cnt["segments"]


def test_reindexed_subtrees_layout():
    """Which Bvh subtrees the device library re-indexes: pure box/primitive subtrees
    become one ACCEL item + an (n_leaves - 1)-node tree; media, frame switches and list elements stay
    on the reference-order stream."""
    def layout(name, bvh, accel=2):
        world, cam = R.build_scene(name, 8, 8, use_bvh=bvh)
        lay = np.zeros(10, np.uint32)
        H.render(world, cam, 8, 8, 1, accel=accel, layout=lay)
        return world.counts(), dict(items=int(lay[0]), nodes=int(lay[1]), accels=int(lay[2]), depth=int(lay[3])), int(lay[5])
    c, l, _ = layout("book1", True)
    assert l["accels"] == 1 and l["nodes"] == c["spheres"] - 2            # one leaf per sphere; the ground sphere stays outside
    assert l["items"] == c["spheres"] + 3 and l["depth"] <= 30            # ACCEL + spheres + the ordered leaf's BBOX + END
    c, l, _ = layout("book1", True, accel=1)                              # fast tree: a leaf's box is derived from its sphere's
    assert l["accels"] == 1 and l["items"] == c["spheres"] + 3            # record, so no BBOX item is kept except the ordered leaf's
    c, l, prisms = layout("book1", True, accel=0)
    assert l == dict(items=c["items"], nodes=0, accels=0, depth=0) and prisms == 0   # the stream as flattened
    c, l, _ = layout("book1", False)
    assert l["accels"] == 0                                               # a plain list has no boxes to re-index
    c, l, prisms = layout("final", False)
    assert l["accels"] == 2 and l["nodes"] == (400 - 1) + (1000 - 1)      # the box floor and the sphere cube
    assert prisms == 400                                                  # every rect_prism: six Rect items -> one record
    assert l["items"] == c["items"] - c["bbox"] + 2 - 5 * 400             # every BBOX item replaced by 2 ACCEL items
    c, l, _ = layout("final", True)                                       # top-level Bvh holds media -> stays threaded,
    assert l["accels"] >= 2 and l["items"] < c["items"]                   # its pure subtrees are still re-indexed
    c, l, prisms = layout("cornell", False)
    assert prisms == 2 and l["items"] == c["items"] - 10                  # the two rotated boxes
    c, l, prisms = layout("cornell_smoke", False, accel=0)
    assert prisms == 2 + 3 and c["media"] == 3                            # smoke boxes + the prisms inside the Bvh-bounded cloud


def test_prism_fusion_and_general_medium_boundaries_match_the_oracle(oracle):
    """rect_prism as one record is the same six Rect::hit calls (object.rs:396-410,420-473): fused and unfused streams,
    all three traversals and the tree-walking oracle agree per sample — also where the prism or a whole Bvh is the
    boundary of a ConstantMedium (object.rs:533-575)."""
    nx, ny, ns = 36, 36, 4
    for name, bvh in (("cornell_smoke", False), ("cornell_smoke", True), ("cornell", True), ("final", False), ("kitchen_sink", True)):
        world, cam = R.build_scene(name, nx, ny, use_bvh=bvh)
        _, osmp, cnt = oracle.Scene(name, nx, ny, top_level_bvh=bvh).render(ns, seed=31, nthreads=4, want_samples=True, want_counters=True)
        for accel in (1, 2, 0):
            for fuse in (True, False):
                _, smp = H.render(world, cam, nx, ny, ns, seed=31, want_samples=True, accel=accel, fuse_prisms=fuse)
                assert n_diff(smp[..., :3], osmp) == 0, (name, bvh, accel, fuse)
                assert int(smp[..., 3].sum()) == cnt["segments"], (name, bvh, accel, fuse)


def test_prism_items_supplied_by_the_host(oracle):
    """RTIOW_ITEM_PRISM is also part of the ABI: a host may send `rect_prism(p0, p1, m)` (object.rs:420-473) as one item
    instead of the six Rect items of its And tree.  The Cornell box with its two prisms rewritten that way by hand renders
    the oracle's bits, in every traversal mode."""
    nx, ny, ns = 40, 40, 5
    world, cam = R.build_scene("cornell", nx, ny, use_bvh=False)
    d = world.desc.contents
    src = (N.Item * d.n_items)()
    C.memmove(src, d.items, C.sizeof(src))
    out, i = [], 0
    while i < d.n_items:
        run = [src[i + k] for k in range(6)] if i + 6 <= d.n_items else []
        axes = [((it.b_w >> 24) >> 2) & 3 for it in run]
        if run and all((it.a_w & 15) == 3 for it in run) and axes == [2, 1, 0, 2, 1, 0] and len({it.a_w >> 4 for it in run}) == 1:
            p = N.Item()                                    # a = p0, b = p1 (see include/rtiow_b200.h)
            p.a[0], p.a[1], p.a[2] = run[5].a[0], run[4].a[0], run[3].a[0]      # k of the x, y, z faces at p0
            p.b[0], p.b[1], p.b[2] = run[2].a[0], run[1].a[0], run[0].a[0]      # k of the x, y, z faces at p1
            p.a_w = 6 | (run[0].a_w & ~15)
            p.b_w = run[0].b_w & 0x00FFFFFF
            out.append(p)
            i += 6
        else:
            out.append(src[i])
            i += 1
    assert len(out) == d.n_items - 10                       # two prisms
    items = (N.Item * len(out))(*out)
    desc = _desc_copy(world)
    desc.items = items
    desc.n_items = len(out)
    assert N.abi().rtiow_b200_scene_validate(C.byref(desc)) == 0, N.abi().rtiow_b200_last_error()
    want, _, _ = oracle.Scene("cornell", nx, ny, top_level_bvh=False).render(ns, nthreads=4)
    for accel in (1, 2, 0):
        img = np.zeros((ny, nx, 3), np.float32)
        rc = H.lib().harness_render(C.byref(desc), C.byref(cam.rec), nx, ny, ns, 0xDEADBEEF, 0, ny, img.ctypes.data, None, accel, None, 1, 1)
        assert rc == 0 and n_diff(img, want) == 0, accel


def test_medium_boundary_validation():
    world, _ = R.build_scene("cornell_smoke", 16, 16, use_bvh=False)
    lib = N.abi()
    items = (N.Item * world.desc.contents.n_items)()

    def fresh():
        C.memmove(items, world.desc.contents.items, C.sizeof(items))
        d = _desc_copy(world)
        d.items = items
        return d

    d = fresh()
    assert lib.rtiow_b200_scene_validate(C.byref(d)) == 0
    media = [i for i in range(len(items)) if (items[i].a_w & 15) == 4]
    assert len(media) == 3
    run_end = lambda i: int(np.float32(items[i].a[2]).view(np.uint32))  # noqa: E731
    assert [run_end(m) - m - 1 for m in media[:2]] == [6, 6]             # a rect_prism boundary: six Rect items
    assert run_end(media[2]) - media[2] - 1 > 9                          # a Bvh boundary: boxes + primitives
    d = fresh()
    items[media[0]].a[2] = float(np.uint32(media[0] + 1).view(np.float32))   # empty boundary
    assert lib.rtiow_b200_scene_validate(C.byref(d)) == N.ERR_INVALID_SCENE
    assert "boundary" in lib.rtiow_b200_last_error().decode()
    d = fresh()
    first_box = next(j for j in range(media[2] + 1, run_end(media[2])) if (items[j].a_w & 15) == 1)
    items[first_box].a_w = 1 | ((run_end(media[2]) + 1) << 4)            # a boundary box that skips out of its run
    assert lib.rtiow_b200_scene_validate(C.byref(d)) == N.ERR_INVALID_SCENE
    d = fresh()
    inner = media[2] + 1
    items[inner].a_w = 4 | (items[inner].a_w & ~15)                      # a medium inside a medium boundary
    assert lib.rtiow_b200_scene_validate(C.byref(d)) == N.ERR_INVALID_SCENE


def test_ordered_leaves_keep_order_dependent_hits(oracle):
    """book-1, pixel (x 192, row 89), sample 48 at 400x200: inside the glass sphere a ray reaches the point where that
    sphere touches the radius-1000 ground sphere; f32 cancellation puts the ground hit (t = 0.7792329) BEFORE the entry of
    the ground's own box (0.779288), with the glass sphere's exit (0.7792501) in between.  The reference asks the ground
    first and keeps it; a nearer-first traversal that lets the later-found glass hit cull the ground's box loses it
    (round 1 did: 1 sample in 4 M on book-1).  The ground sphere is therefore an "ordered leaf" (scene_blob.hpp) and
    every traversal gives the oracle's bits."""
    nx, ny, ns = 400, 200, 50
    world, cam = R.build_scene("book1", nx, ny)
    lay = np.zeros(10, np.uint32)
    H.render(world, cam, 8, 8, 1, accel=1, layout=lay)
    assert lay[7] == 1                                                   # exactly one ordered leaf: the ground
    want, osmp, _ = oracle.Scene("book1", nx, ny).render(ns, nthreads=4, rows=(89, 90), want_samples=True)
    for accel in (1, 2, 0):
        got, smp = H.render(world, cam, nx, ny, ns, accel=accel, rows=(89, 90), want_samples=True)
        assert n_diff(smp[..., :3], osmp) == 0 and n_diff(got, want) == 0, accel
        assert smp[0, 192, 48, 3] == 17                                  # the path that used to run 25 segments


def test_print_ppm_formatting(tmp_path, oracle):
    rgb = np.array([[[0.25, 1.0, 0.0], [4.0, -1.0, np.nan]]], np.float32)
    path = tmp_path / "a.ppm"
    R.print_ppm(R.Image(rgb), path)
    assert path.read_text() == "P3\n2 1\n255\n127 255 0\n255 0 0\n"      # lib.rs:345,358: one pixel per line
    assert oracle.ppm_quantise(rgb).reshape(-1).tolist() == [127, 255, 0, 255, 0, 0]


def _adversarial_rays(rng, n, lo, hi, planes):
    """Rays that stress the box tests: random, axis-parallel (exact +0/-0 components), origins exactly on
    box planes of the scene, tiny and huge direction components, denormals."""
    o = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    k = np.arange(n)
    special = [0.0, -0.0, 1e-38, -1e-38, 1e-45, 1e30, -1e30, 1e-20, 3e38, np.inf]
    for j, v in enumerate(special):                     # one component special
        sel = k % 23 == j
        d[sel, (k[sel] // 23) % 3] = v
    sel = k % 23 == 11                                  # two components zero: axis-parallel
    ax = (k[sel] // 23) % 3
    d[sel] = 0.0
    d[sel, ax] = np.where(k[sel] % 2 == 0, 1.0, -1.0)
    for cls, zero_dir in ((12, False), (13, True)):     # origin exactly on a box plane (0 * inf = NaN in Aabb::hit)
        sel = np.where(k % 23 == cls)[0]
        ax = (sel // 23) % 3
        o[sel, ax] = planes[rng.integers(0, len(planes), size=len(sel)), ax]
        if zero_dir:
            d[sel, ax] = np.where(sel % 2 == 0, 0.0, -0.0)
    sel = k % 23 == 14                                  # huge origin
    o[sel] *= 1e9
    t = rng.uniform(0, 1, size=(n, 1)).astype(np.float32)
    return np.concatenate([o, d, t], axis=1)


@pytest.mark.parametrize("name,bvh,lo,hi", [("book1", True, -12, 12), ("bench_cornell", True, -50, 600),
                                            ("final", False, -300, 700), ("kitchen_sink", True, -10, 10)])
def test_traversals_agree_on_adversarial_rays(name, bvh, lo, hi):
    """The conservative inner-box test (traversal 0) must never lose a hit the reference's Aabb::hit keeps:
    same winner and same t, bit for bit, as the exact re-indexed tree (2) and the reference-order stream (0)
    on rays built to hit the corner cases (zero / denormal / huge direction components, origins on planes)."""
    world, _ = R.build_scene(name, 16, 16, use_bvh=bvh)
    items = world.items()
    boxes = items[(items[:, 3] & 15) == 1]
    planes = np.concatenate([boxes[:, 0:3], boxes[:, 4:7]]).view(np.float32)
    rays = _adversarial_rays(np.random.default_rng(7), 60000, lo, hi, planes)
    with np.errstate(all="ignore"):
        ref = H.trace_rays(world, rays, accel=0)
        exact = H.trace_rays(world, rays, accel=2)
        fast = H.trace_rays(world, rays, accel=1)
    assert (ref[:, 0] != 0xffffffff).sum() > 1000                      # the test does hit things
    # A ray lying exactly in a Rect's plane gives t = 0/0 = NaN in the reference too (object.rs:194-197 lets NaN
    # through both range checks); which hit survives then depends on the sibling order of the reference's own
    # tree (bvh.rs:104-111), so only the reference-order stream can follow it.  Such rays (a zero direction
    # component in a scene with Rects) are only required to terminate; everything else must agree bit for bit.
    has_rects = ((items[:, 3] & 15) == 3).any()
    ok = ~((rays[:, 3:6] == 0).any(axis=1)) if has_rects else np.ones(len(rays), bool)
    assert ok.sum() > 45000
    assert np.array_equal(exact[ok], ref[ok])
    assert np.array_equal(fast[ok], ref[ok])


def test_conservative_box_test_error_bound():
    """the inner-node test uses n = fma(plane, 1/d, -fl(o * 1/d)) instead of the reference's
    fl(fl(plane - o) * 1/d) (aabb.rs:19-21) and accepts with a slack of 16 * 2^-24 * (|plane| + |o|) * |1/d| per
    side.  Analysis bounds the difference by 4 of those units; measure it on 2e7 draws across 12 orders of
    magnitude, a third of them with plane ~ o (catastrophic cancellation)."""
    rng = np.random.default_rng(1)
    worst = 0.0
    for _ in range(5):
        n = 4_000_000
        scale = np.float32(10.0) ** rng.uniform(-6, 6, size=n).astype(np.float32)
        p = rng.normal(size=n).astype(np.float32) * scale
        near = p * (np.float32(1) + rng.normal(size=n).astype(np.float32) * np.float32(1e-6))
        far = rng.normal(size=n).astype(np.float32) * scale * np.float32(10.0) ** rng.integers(-3, 4, size=n).astype(np.float32)
        o = np.where(rng.integers(0, 3, size=n) == 0, near, far).astype(np.float32)
        d = rng.normal(size=n).astype(np.float32) * np.float32(10.0) ** rng.uniform(-8, 3, size=n).astype(np.float32)
        with np.errstate(all="ignore"):
            i = (np.float32(1) / d).astype(np.float32)
            ref = ((p - o).astype(np.float32) * i).astype(np.float32)                       # the reference's arithmetic
            c = (-(o * i).astype(np.float32)).astype(np.float32)
            fast = (p.astype(np.float64) * i.astype(np.float64) + c.astype(np.float64)).astype(np.float32)  # one rounding = fma
            unit = (np.abs(p).astype(np.float64) + np.abs(o).astype(np.float64)) * np.abs(i).astype(np.float64) * 2.0 ** -24
            ok = (np.abs(i) > 2.0 ** -100) & (np.abs(i) < 2.0 ** 100) & np.isfinite(ref) & np.isfinite(fast) & (unit > 1e-300) & (unit < 1e30)
            worst = max(worst, float((np.abs(ref.astype(np.float64) - fast.astype(np.float64))[ok] / unit[ok]).max()))
    assert worst <= 4.0, worst      # the analytic bound
    assert worst * 4 <= 16.0        # the kernel's slack per side leaves a factor of 4


def test_feature_specialisations_match_the_general_code(oracle):
    """The feature-specialised instantiations of the per-path code (SceneT<Mem, kFeat>: everything the scene
    cannot contain compiled out) are picked for book-1 (spheres) and the Cornell box (rect list) only, and give
    the same bits as the general code."""
    nx, ny, ns = 48, 32, 6
    for name, bvh, profile in (("book1", True, 1), ("book1", False, 1), ("cornell", False, 2), ("cornell_empty", False, 2),
                               ("bench_cornell", True, 3), ("kitchen_sink", True, 0), ("final", False, 3), ("final", True, 3),
                               ("volume_test", False, 3), ("cornell_smoke", False, 0), ("simple_light", True, 0)):   # simple_light: an ordered leaf (the radius-1000 light)
        world, cam = R.build_scene(name, nx, ny, use_bvh=bvh)
        lay = np.zeros(10, np.uint32)
        general, gs = H.render(world, cam, nx, ny, ns, want_samples=True, accel=1)
        special, ss = H.render(world, cam, nx, ny, ns, want_samples=True, accel=1, lean=True, layout=lay)
        assert int(lay[4]) == profile, (name, int(lay[4]))
        assert np.array_equal(general.view(np.uint32), special.view(np.uint32)) and np.array_equal(gs.view(np.uint32), ss.view(np.uint32))
    for name, bvh in (("book1", True), ("cornell", False)):
        want, _, _ = oracle.Scene(name, nx, ny, top_level_bvh=bvh).render(ns, nthreads=os.cpu_count() or 4)
        world, cam = R.build_scene(name, nx, ny, use_bvh=bvh)
        for accel in (1, 2, 0):
            got, _ = H.render(world, cam, nx, ny, ns, accel=accel, lean=True)
            assert n_diff(got, want) == 0, (name, accel)


@pytest.mark.parametrize("nx,n_rows,s_count,world,open_scene,warps,max_strips,forced", [
    (1200, 800, 50, 1, True, 4736, 4096, 0),      # C2 on one GPU: 37500 tiles in strips of 16, units of 8 samples
    (1200, 800, 50, 8, True, 4736, 4096, 0),      # C2 on 8 GPUs: every rank takes every 8th tile, units of 2 samples
    (403, 301, 7, 3, True, 4736, 4096, 0),        # ragged on both edges, a world size that divides nothing
    (800, 800, 37, 4, False, 3552, 4096, 0),      # closed scene: two unit sizes (big chunks + a tail of smaller ones)
    (333, 97, 5, 2, False, 96, 64, 0),            # few warps and a small order table: strips of several tiles, the last one padded
    (64, 48, 9, 5, True, 4736, 1, 4),             # one strip for everything (no room for a table), forced chunk, ragged last chunk
])
def test_work_units_cover_every_pixel_sample_exactly_once(nx, n_rows, s_count, world, open_scene, warps, max_strips, forced):
    """The multi-GPU partition and the unit scheduling, walked on the CPU with the very functions the megakernel's refill
    and the fold use (path_logic.cuh unit_samples / tile_of_rank / tile_origin / TileMap) and the host's own plan
    (unit_plan.hpp): over the ranks of a world, with a shuffled strip order per rank (any order must do: it is a hint),
    every pixel-sample of the frame is handed out exactly once, and every staging slot folds back to the pixel that was
    rendered into it."""
    rng = np.random.default_rng(nx * 131 + world)
    total = np.zeros((n_rows, nx, s_count), np.uint16)
    for rank in range(world):
        _, _, plan = H.unit_coverage(nx, n_rows, 1, rank, world, forced, open_scene, warps, max_strips)
        order = rng.permutation(plan["n_strips"]).astype(np.uint32)
        counts, bad, plan = H.unit_coverage(nx, n_rows, s_count, rank, world, forced, open_scene, warps, max_strips, order=order)
        assert bad == 0, (rank, plan)
        assert plan["n_strips"] <= max(1, max_strips) and (plan["n_strips"] << plan["order_shift"]) >= plan["n_groups"]
        assert counts.max() <= 1, (rank, plan)                      # nothing twice within a rank
        total += counts
        default, bad, _ = H.unit_coverage(nx, n_rows, s_count, rank, world, forced, open_scene, warps, max_strips)
        assert bad == 0 and np.array_equal(default, counts)        # the order changes WHEN a tile is rendered, never WHETHER
    assert total.min() == 1 and total.max() == 1                    # every pixel-sample exactly once over the world
    # the ranks' shares are spread evenly: tile counts differ by at most one
    tiles = ((nx + 7) // 8) * ((n_rows + 3) // 4)
    shares = [H.unit_coverage(nx, n_rows, 1, r, world, forced, open_scene, warps, max_strips)[2]["n_groups"] for r in range(world)]
    assert sum(shares) == tiles and max(shares) - min(shares) <= 1


def test_unit_size_rule_picks_the_measured_optima():
    """unit_plan.hpp: open scenes get one unit size, the largest c of 8, 4, 2, 1 with 4 c^2 <= one-sample units per resident
    warp — the sizes measured best for the whole book-1 frame and a half, a quarter and an eighth of it on 148 x 32 warps
    (profiles/r02/p1_unit_size/); closed scenes get big chunks plus a tail of smaller ones that covers every sample."""
    warps = 148 * 32
    for world, want in ((1, 8), (2, 4), (4, 4), (8, 2)):
        plan = H.unit_coverage(1200, 800, 50, 0, world, 0, True, warps, 4096)[2]
        assert plan["s_chunk"] == want and plan["s_tail_begin"] == 50, (world, plan)
    plan = H.unit_coverage(800, 800, 100, 0, 1, 0, False, 148 * 32, 4096)[2]       # Cornell at 100 spp on the 1024-thread kernel
    assert plan["s_chunk"] == 8 and plan["s_chunk_tail"] == 4 and plan["s_tail_begin"] == 80, plan
    assert H.unit_coverage(64, 48, 3, 0, 1, 0, True, warps, 4096)[2]["s_chunk"] == 1     # a tiny frame: single samples
