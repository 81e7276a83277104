"""Host-side mirror of the reference crate's surface for the render path, over the C ABI.

Reference (cbiffle/rtiow-rust)                        here
------------------------------------------------      -----------------------------------------
camera::Camera::look(...)      src/camera.rs:18       Camera.look(...)
par_cast(nx, ny, ns, &camera, world)  src/lib.rs:363  par_cast(nx, ny, ns, camera, world, seed=...)
cast(nx, ny, ns, &camera, world, &mut rng)  :378      cast(nx, ny, ns, camera, world, seed=...)
print_ppm(image)               src/lib.rs:344         print_ppm(image, file)
cornell_box_with_boxes(), book_final_scene(), ...     build_scene(name, nx, ny, scene_seed, use_bvh)
Image(Vec<Vec<Vec3>>)          src/lib.rs:321         Image (numpy [ny, nx, 3] float32, row 0 = top)

The scene objects themselves (Sphere, Rect, Translate, Bvh ...) are built by the C++ host mirror
(csrc/host, namespace rtiow::object) — Python only holds handles.  Everything below par_cast runs in
the sm_100a megakernel; a missing extension or GPU raises, nothing falls back to the CPU.
"""
import ctypes as C
import sys

import numpy as np

from . import _native as N

__all__ = ["Camera", "World", "Image", "RtiowError", "build_scene", "par_cast", "par_cast_multi", "par_cast_ppm", "cast", "print_ppm", "ppm_bytes", "ppm_bytes_device",
           "SCENES", "DEFAULT_SEED"]

DEFAULT_SEED = 0xDEADBEEF  # src/main.rs:333
SCENES = ("book1", "book1_head", "cornell", "cornell_empty", "bench_cornell", "final", "motion_test", "volume_test",
          "simple_light", "kitchen_sink", "cornell_smoke")


class RtiowError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"[rtiow_b200 error {code}] {message}")
        self.code = code


def _check(rc, lib=None):
    if rc != 0:
        raise RtiowError(rc, (lib or N.abi()).rtiow_b200_last_error().decode())


class Camera:
    """src/camera.rs:6-15.  Wraps the #[repr(C)] export the ABI consumes."""

    def __init__(self, rec):
        self.rec = rec

    @staticmethod
    def look(look_from, look_at, up, fov, aspect, aperture, focus_dist, exposure=(0.0, 1.0)):
        """Camera::look (src/camera.rs:18-50); computed by the C++ host mirror in f32."""
        rec = N.CameraRec()
        f3 = lambda v: (C.c_float * 3)(*[float(x) for x in v])  # noqa: E731
        if N.host().rtiow_host_camera_look(f3(look_from), f3(look_at), f3(up), fov, aspect, aperture, focus_dist,
                                           exposure[0], exposure[1], C.byref(rec)):
            raise RuntimeError(N.host().rtiow_host_last_error().decode())
        return Camera(rec)

    def as_array(self):
        return np.frombuffer(bytes(self.rec), np.float32).copy()


class Image:
    """lib.rs:321 — rows top first; `.rgb` is linear (pre-gamma) float32 [ny, nx, 3]."""

    def __init__(self, rgb):
        self.rgb = rgb

    @property
    def ny(self):
        return self.rgb.shape[0]

    @property
    def nx(self):
        return self.rgb.shape[1]


class World:
    """`impl World` (src/lib.rs:23-55): a flattened scene (list or Bvh top level) plus, lazily, its
    device-resident copy.  Obtain one from build_scene()."""

    def __init__(self, host_handle, name, flavour="parity"):
        self._h = host_handle
        self.name = name
        self.flavour = flavour      # which build of the device library renders it: "parity" (default) or "fast"
        self._gpu = {}
        self._scene_bytes = {}

    @property
    def lib(self):
        return N.abi(self.flavour)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def close(self):
        for h in list(self._gpu.values()):
            self.lib.rtiow_b200_scene_destroy(h)
        self._gpu = {}
        if self._h:
            N.host().rtiow_host_scene_free(self._h)
            self._h = None

    def __len__(self):
        return N.host().rtiow_host_scene_len(self._h)

    @property
    def desc(self):
        return N.host().rtiow_host_scene_desc(self._h)

    def validate(self):
        _check(self.lib.rtiow_b200_scene_validate(self.desc), self.lib)

    def items(self):
        d = self.desc.contents
        raw = np.ctypeslib.as_array(C.cast(d.items, C.POINTER(C.c_uint32)), shape=(d.n_items, 8)).copy()
        return raw

    def counts(self):
        d = self.desc.contents
        kinds = self.items()[:, 3] & 15
        return dict(items=d.n_items, frames=d.n_frames, ops=d.n_ops, materials=d.n_materials, textures=d.n_textures,
                    bbox=int((kinds == 1).sum()), spheres=int((kinds == 2).sum()), rects=int((kinds == 3).sum()),
                    media=int((kinds == 4).sum()), set_frames=int((kinds == 5).sum()),
                    background=d.background_kind)

    def gpu(self, device=0):
        """rtiow_b200_scene_create: upload to `device` (once) and return the handle."""
        if device not in self._gpu:
            h = C.c_void_p()
            _check(self.lib.rtiow_b200_scene_create(self.desc, device, C.byref(h)), self.lib)
            self._gpu[device] = h
        return self._gpu[device]

    def upload_fresh(self, device=0):
        """Drop and re-create the device copy (used by the end-to-end timing: H2D inside the timed region)."""
        if device in self._gpu:
            self.lib.rtiow_b200_scene_destroy(self._gpu.pop(device))
        return self.gpu(device)

    def set_tuning(self, device=0, cta_threads=0, ctas_per_sm=0, staging_mib=0, force_global=False):
        _check(self.lib.rtiow_b200_set_tuning(self.gpu(device), cta_threads, ctas_per_sm, staging_mib, int(force_global)), self.lib)

    def set_specialisation(self, enable=True, device=0):
        """False forces the general megakernel even if the scene qualifies for a specialised one.  Same image."""
        _check(self.lib.rtiow_b200_set_specialisation(self.gpu(device), int(bool(enable))), self.lib)

    def set_traversal(self, mode, device=0):
        """0 = re-indexed Bvh subtrees, conservative inner box test (default); 1 = the reference's own visiting
        order; 2 = re-indexed with the reference's box test at every node.  Same image every way."""
        _check(self.lib.rtiow_b200_set_traversal(self.gpu(device), int(mode)), self.lib)

    def scene_bytes(self, device=0):
        """Bytes of the device image of the scene the last render used (uploaded by scene_create).  Asked once per
        device (rtiow_b200_get_stats synchronises the device); the image of a re-uploaded scene has the same size."""
        if device not in self._scene_bytes:
            self._scene_bytes[device] = self.stats(device)["scene_bytes"]
        return self._scene_bytes[device]

    def stats(self, device=0):
        st = N.Stats()
        _check(self.lib.rtiow_b200_get_stats(self.gpu(device), C.byref(st)), self.lib)
        return {f: getattr(st, f) for f, _ in st._fields_}


def build_scene(name, nx, ny, scene_seed=DEFAULT_SEED, use_bvh=True, flavour="parity"):
    """The reference's scene functions by name (src/lib.rs:103-193,237-319; src/main.rs:10-319;
    benches/scene.rs).  Returns (world, camera).  use_bvh mirrors USE_BVH (src/main.rs:321).
    flavour="fast" renders it with the tolerance build of the device library (make FAST=1)."""
    h = N.host().rtiow_host_scene_build(name.encode(), nx, ny, scene_seed, int(use_bvh))
    if not h:
        raise RuntimeError(N.host().rtiow_host_last_error().decode())
    world = World(h, name, flavour)
    cam = N.CameraRec()
    C.memmove(C.byref(cam), N.host().rtiow_host_scene_camera(h), C.sizeof(cam))
    return world, Camera(cam)


def par_cast(nx, ny, ns, camera, world, seed=DEFAULT_SEED, device=0, rows=None):
    """par_cast (src/lib.rs:363-376) on the GPU.  Host buffers in, host Image out.
    `rows=(begin, end)` renders only those output rows (row 0 = top)."""
    r0, r1 = rows if rows is not None else (0, ny)
    out = np.empty((r1 - r0, nx, 3), np.float32)
    _check(world.lib.rtiow_b200_render_rows(world.gpu(device), C.byref(camera.rec), nx, ny, ns, seed, r0, r1,
                                          out.ctypes.data), world.lib)
    return Image(out)


def par_cast_multi(nx, ny, ns, camera, worlds, seed=DEFAULT_SEED):
    """par_cast over several GPUs from ONE host thread (rtiow_b200_render_multi): `worlds` = the same scene built once per
    device, in device order 0, 1, ...  The frame's 8x4-pixel tiles are dealt round-robin, every GPU's fold stores its tiles
    straight into GPU 0's frame over NVLink, GPU 0 copies the frame out.  Bit-identical to par_cast."""
    lib = worlds[0].lib
    handles = (C.c_void_p * len(worlds))(*[w.gpu(i) for i, w in enumerate(worlds)])
    out = np.empty((ny, nx, 3), np.float32)
    _check(lib.rtiow_b200_render_multi(handles, len(worlds), C.byref(camera.rec), nx, ny, ns, seed, out.ctypes.data), lib)
    return Image(out)


def cast(nx, ny, ns, camera, world, seed=DEFAULT_SEED, device=0):
    """cast (src/lib.rs:378-397).  The reference's sequential twin exists to be deterministic; here
    determinism comes from the per-(pixel, sample) RNG key, so it is the same computation."""
    return par_cast(nx, ny, ns, camera, world, seed=seed, device=device)


def render_samples(nx, ny, ns, camera, world, seed=DEFAULT_SEED, device=0, rows=None):
    """Per-sample radiance before the fold: [rows, nx, ns, 4] = r, g, b, path segments."""
    r0, r1 = rows if rows is not None else (0, ny)
    out = np.empty((r1 - r0, nx, ns, 4), np.float32)
    _check(world.lib.rtiow_b200_render_samples(world.gpu(device), C.byref(camera.rec), nx, ny, ns, seed, r0, r1,
                                             out.ctypes.data), world.lib)
    return out


def render_rows_device(nx, ny, ns, camera, world, out_tensor, rows, seed=DEFAULT_SEED, stream=None, row_step=1, row_band=1):
    """Enqueue bands of row_band rows starting at begin, begin + row_step, ... (clipped to end, packed) into
    a CUDA torch tensor (float32) on `stream` (a torch.cuda.Stream, default: current).  Nothing is synchronised."""
    import torch
    assert out_tensor.is_cuda and out_tensor.dtype == torch.float32 and out_tensor.is_contiguous()
    r0, r1 = rows
    assert out_tensor.numel() >= sum(min(row_band, r1 - b) for b in range(r0, r1, row_step)) * nx * 3
    dev = out_tensor.device.index or 0
    s = stream if stream is not None else torch.cuda.current_stream(dev)
    _check(world.lib.rtiow_b200_render_rows_strided_device(world.gpu(dev), C.byref(camera.rec), nx, ny, ns, seed, r0, r1, row_step, row_band,
                                                         C.c_void_p(out_tensor.data_ptr()), C.c_void_p(s.cuda_stream)), world.lib)


def ppm_bytes(image, world=None, device=0):
    """print_ppm's quantiser (sqrt, then to_u8; src/lib.rs:344-361) on the device -> uint8 [ny, nx, 3]."""
    rgb = np.ascontiguousarray(image.rgb if isinstance(image, Image) else image, np.float32)
    out = np.empty(rgb.shape, np.uint8)
    _check(world.lib.rtiow_b200_ppm_quantise(world.gpu(device), rgb.ctypes.data, rgb.size, out.ctypes.data), world.lib)
    return out


def ppm_bytes_device(frame, world, stream=None):
    """The same quantiser on a float32 CUDA tensor, result a uint8 CUDA tensor of the same shape; enqueued on `stream`
    (default: current), nothing synchronised — the float frame never crosses PCIe."""
    import torch
    assert frame.is_cuda and frame.dtype == torch.float32 and frame.is_contiguous()
    dev = frame.device.index or 0
    out = torch.empty(frame.shape, dtype=torch.uint8, device=frame.device)
    s = stream if stream is not None else torch.cuda.current_stream(dev)
    _check(world.lib.rtiow_b200_ppm_quantise_device(world.gpu(dev), C.c_void_p(frame.data_ptr()), frame.numel(),
                                                  C.c_void_p(out.data_ptr()), C.c_void_p(s.cuda_stream)), world.lib)
    return out


def par_cast_ppm(nx, ny, ns, camera, world, seed=DEFAULT_SEED, device=0):
    """par_cast + print_ppm's quantiser in one device-side call (rtiow_b200_render_ppm): uint8 [ny, nx, 3], the numbers
    print_ppm writes (src/lib.rs:346-360); a quarter of the float frame's bytes come back over PCIe."""
    out = np.empty((ny, nx, 3), np.uint8)
    _check(world.lib.rtiow_b200_render_ppm(world.gpu(device), C.byref(camera.rec), nx, ny, ns, seed, out.ctypes.data), world.lib)
    return out


def print_ppm(image, path_or_file=None):
    """print_ppm (src/lib.rs:344-361): ASCII P3, one pixel per line, byte-identical formatting."""
    rgb = np.ascontiguousarray(image.rgb, np.float32)
    if path_or_file is None or path_or_file is sys.stdout:
        path_or_file = "/dev/stdout"
    if N.host().rtiow_host_print_ppm(rgb.ctypes.data, image.nx, image.ny, str(path_or_file).encode()):
        raise RuntimeError(N.host().rtiow_host_last_error().decode())
