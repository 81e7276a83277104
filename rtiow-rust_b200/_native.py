"""ctypes declarations for the two in-tree native libraries:

  _build/librtiow_b200.so — the C ABI of include/rtiow_b200.h (sm_100a kernels inside)
  _build/librtiow_host.so — extern "C" facade over the C++ host mirror of the crate (csrc/host)

There is no pure-Python or CPU implementation behind these: if a library is missing the import
fails loudly with the command that builds it.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RTIOW_B200_BUILD_DIR selects another in-tree build (`make OUT=_build_x`): same-box A/B timing of two kernels
BUILD_DIR = os.path.join(_HERE, os.environ.get("RTIOW_B200_BUILD_DIR", "_build"))
ABI_LIB = os.path.join(BUILD_DIR, "librtiow_b200.so")
# the tolerance build of the same library (make FAST=1): FMA contraction, approximate division / square root
FAST_ABI_LIB = os.path.join(_HERE, os.environ.get("RTIOW_B200_FAST_BUILD_DIR", "_build_fast"), "librtiow_b200.so")
HOST_LIB = os.path.join(BUILD_DIR, "librtiow_host.so")

PEER_HANDLE_BYTES = 128
RTIOW_OK, ERR_INVALID_ARG, ERR_INVALID_SCENE, ERR_UNSUPPORTED, ERR_CUDA, ERR_NO_DEVICE = range(6)


class Item(C.Structure):
    _fields_ = [("a", C.c_float * 3), ("a_w", C.c_uint32), ("b", C.c_float * 3), ("b_w", C.c_uint32)]


class XformOp(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("v", C.c_float * 3)]


class Frame(C.Structure):
    _fields_ = [("first_op", C.c_uint32), ("n_ops", C.c_uint32)]


class MaterialRec(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("tex", C.c_uint32), ("albedo", C.c_float * 3), ("param", C.c_float),
                ("reserved", C.c_uint32 * 2)]


class TextureRec(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("color", C.c_float * 3), ("scale", C.c_float), ("child0", C.c_uint32),
                ("child1", C.c_uint32), ("reserved", C.c_uint32)]


class SceneDesc(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("n_items", C.c_uint32), ("items", C.POINTER(Item)),
                ("n_frames", C.c_uint32), ("n_ops", C.c_uint32), ("frames", C.POINTER(Frame)),
                ("ops", C.POINTER(XformOp)), ("n_materials", C.c_uint32), ("n_textures", C.c_uint32),
                ("materials", C.POINTER(MaterialRec)), ("textures", C.POINTER(TextureRec)),
                ("perlin_vecs", C.POINTER(C.c_float)), ("perlin_perm", C.POINTER(C.c_uint8)),
                ("background_kind", C.c_uint32), ("background_c0", C.c_float * 3), ("background_c1", C.c_float * 3)]


class CameraRec(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("lower_left_corner", C.c_float * 3), ("horizontal", C.c_float * 3),
                ("vertical", C.c_float * 3), ("u", C.c_float * 3), ("v", C.c_float * 3), ("lens_radius", C.c_float),
                ("time0", C.c_float), ("time1", C.c_float)]


class Stats(C.Structure):
    _fields_ = [("trace_ms", C.c_double), ("reduce_ms", C.c_double), ("samples", C.c_uint64),
                ("segments", C.c_uint64), ("kernel_launches", C.c_uint32), ("passes", C.c_uint32),
                ("scene_in_smem", C.c_uint32), ("scene_bytes", C.c_uint32), ("grid", C.c_uint32),
                ("block", C.c_uint32), ("dyn_smem_bytes", C.c_uint32), ("regs_per_thread", C.c_uint32),
                ("accel_nodes", C.c_uint32), ("accel_subtrees", C.c_uint32), ("traversal", C.c_uint32), ("kernel_profile", C.c_uint32)]


# every symbol include/rtiow_b200.h declares
ABI_SYMBOLS = ("rtiow_b200_abi_version", "rtiow_b200_build_flavour", "rtiow_b200_last_error", "rtiow_b200_scene_validate",
               "rtiow_b200_scene_create", "rtiow_b200_scene_destroy", "rtiow_b200_release_cached_memory", "rtiow_b200_render", "rtiow_b200_render_rows",
               "rtiow_b200_render_rows_device", "rtiow_b200_render_rows_strided_device", "rtiow_b200_render_samples", "rtiow_b200_ppm_quantise",
               "rtiow_b200_ppm_quantise_device", "rtiow_b200_render_ppm",
               "rtiow_b200_render_multi", "rtiow_b200_peer_frame_create", "rtiow_b200_peer_frame_export", "rtiow_b200_peer_frame_connect",
               "rtiow_b200_peer_frame_ptr", "rtiow_b200_peer_frame_destroy", "rtiow_b200_render_peers",
               "rtiow_b200_get_stats", "rtiow_b200_set_tuning", "rtiow_b200_set_traversal", "rtiow_b200_set_specialisation")

_abi = {}
_host = None


def _missing(path):
    return ImportError(
        f"{path} is missing: the native CUDA path is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
        f"(or `make -C {_HERE}`) — there is no Python/CPU fallback for the render path.")


def _declare(L):
    u32, u64, vp = C.c_uint32, C.c_uint64, C.c_void_p
    L.rtiow_b200_abi_version.restype = C.c_int
    L.rtiow_b200_build_flavour.restype = C.c_char_p
    L.rtiow_b200_last_error.restype = C.c_char_p
    L.rtiow_b200_scene_validate.argtypes = [C.POINTER(SceneDesc)]
    L.rtiow_b200_scene_create.argtypes = [C.POINTER(SceneDesc), C.c_int, C.POINTER(vp)]
    L.rtiow_b200_scene_destroy.argtypes = [vp]
    L.rtiow_b200_scene_destroy.restype = None
    L.rtiow_b200_release_cached_memory.restype = None
    L.rtiow_b200_render.argtypes = [vp, C.POINTER(CameraRec), u32, u32, u32, u64, vp]
    L.rtiow_b200_render_rows.argtypes = [vp, C.POINTER(CameraRec), u32, u32, u32, u64, u32, u32, vp]
    L.rtiow_b200_render_rows_device.argtypes = [vp, C.POINTER(CameraRec), u32, u32, u32, u64, u32, u32, vp, vp]
    L.rtiow_b200_render_rows_strided_device.argtypes = [vp, C.POINTER(CameraRec), u32, u32, u32, u64, u32, u32, u32, u32, vp, vp]
    L.rtiow_b200_render_samples.argtypes = [vp, C.POINTER(CameraRec), u32, u32, u32, u64, u32, u32, vp]
    L.rtiow_b200_ppm_quantise.argtypes = [vp, vp, C.c_size_t, vp]
    L.rtiow_b200_ppm_quantise_device.argtypes = [vp, vp, C.c_size_t, vp, vp]
    L.rtiow_b200_render_ppm.argtypes = [vp, C.POINTER(CameraRec), u32, u32, u32, u64, vp]
    L.rtiow_b200_render_multi.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(CameraRec), u32, u32, u32, u64, vp]
    L.rtiow_b200_peer_frame_create.argtypes = [C.c_int, u32, u32, u32, u32, C.POINTER(vp)]
    L.rtiow_b200_peer_frame_export.argtypes = [vp, vp]
    L.rtiow_b200_peer_frame_connect.argtypes = [vp, vp]
    L.rtiow_b200_peer_frame_ptr.argtypes = [vp, C.POINTER(vp)]
    L.rtiow_b200_peer_frame_destroy.argtypes = [vp]
    L.rtiow_b200_peer_frame_destroy.restype = None
    L.rtiow_b200_render_peers.argtypes = [vp, C.POINTER(CameraRec), u32, u32, u32, u64, vp, vp]
    L.rtiow_b200_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.rtiow_b200_set_tuning.argtypes = [vp, u32, u32, u32, C.c_int]
    L.rtiow_b200_set_traversal.argtypes = [vp, C.c_int]
    L.rtiow_b200_set_specialisation.argtypes = [vp, C.c_int]
    return L


def abi(flavour="parity"):
    """The C ABI library.  flavour "parity" (default; what every parity test pins) or "fast" (the tolerance build,
    `make FAST=1`): the same entry points, loaded side by side (hidden internals, -Bsymbolic)."""
    if flavour not in _abi:
        if flavour == "parity":
            if not os.path.exists(ABI_LIB):
                raise _missing(ABI_LIB)
            _abi[flavour] = _declare(C.CDLL(ABI_LIB, mode=C.RTLD_GLOBAL))
        elif flavour == "fast":
            if not os.path.exists(FAST_ABI_LIB):
                raise _missing(FAST_ABI_LIB)
            _abi[flavour] = _declare(C.CDLL(FAST_ABI_LIB))
            assert _abi[flavour].rtiow_b200_build_flavour().startswith(b"fast")
        else:
            raise ValueError(f"unknown library flavour {flavour!r}")
    return _abi[flavour]


def host():
    global _host
    if _host is None:
        abi()  # librtiow_host.so links against librtiow_b200.so
        if not os.path.exists(HOST_LIB):
            raise _missing(HOST_LIB)
        L = C.CDLL(HOST_LIB)
        u32, u64, vp, fp = C.c_uint32, C.c_uint64, C.c_void_p, C.POINTER(C.c_float)
        L.rtiow_host_last_error.restype = C.c_char_p
        L.rtiow_host_scene_build.restype = vp
        L.rtiow_host_scene_build.argtypes = [C.c_char_p, u32, u32, u64, C.c_int]
        L.rtiow_host_scene_free.argtypes = [vp]
        L.rtiow_host_scene_free.restype = None
        L.rtiow_host_scene_desc.restype = C.POINTER(SceneDesc)
        L.rtiow_host_scene_desc.argtypes = [vp]
        L.rtiow_host_scene_camera.restype = C.POINTER(CameraRec)
        L.rtiow_host_scene_camera.argtypes = [vp]
        L.rtiow_host_scene_len.restype = u32
        L.rtiow_host_scene_len.argtypes = [vp]
        L.rtiow_host_camera_look.argtypes = [fp, fp, fp] + [C.c_float] * 6 + [C.POINTER(CameraRec)]
        L.rtiow_host_print_ppm.argtypes = [vp, u32, u32, C.c_char_p]
        L.rtiow_host_par_cast.argtypes = [vp, u32, u32, u32, u64, C.c_int, vp]
        _host = L
    return _host
