//! UNCOMPILED (see ../../README.md).  Raw bindings to `include/rtiow_b200.h`, ABI version 2.
//! Every struct is the `#[repr(C)]` twin of the C struct of the same name; every function the
//! `extern "C"` twin of the C prototype.  Sizes are pinned by the const assertions at the bottom
//! (the C side pins the same numbers in `tests/test_host_cpu.py::test_abi_struct_sizes_match_header`).
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_int, c_void};

pub const RTIOW_B200_ABI_VERSION: u32 = 2;

pub const RTIOW_OK: c_int = 0;
pub const RTIOW_ERR_INVALID_ARG: c_int = 1;
pub const RTIOW_ERR_INVALID_SCENE: c_int = 2;
pub const RTIOW_ERR_UNSUPPORTED: c_int = 3;
pub const RTIOW_ERR_CUDA: c_int = 4;
pub const RTIOW_ERR_NO_DEVICE: c_int = 5;

// item kinds: word a_w = kind | (payload << 4); word b_w = material | (flags << 24)
pub const RTIOW_ITEM_END: u32 = 0;
pub const RTIOW_ITEM_BBOX: u32 = 1;
pub const RTIOW_ITEM_SPHERE: u32 = 2;
pub const RTIOW_ITEM_RECT: u32 = 3;
pub const RTIOW_ITEM_MEDIUM: u32 = 4;
pub const RTIOW_ITEM_SET_FRAME: u32 = 5;
pub const RTIOW_ITEM_PRISM: u32 = 6;

pub const RTIOW_FLAG_HAS_OFFSET: u32 = 1;
pub const RTIOW_FLAG_FLIP: u32 = 2;
pub const RTIOW_FLAG_AXIS_SHIFT: u32 = 2;

pub const RTIOW_OP_TRANSLATE: u32 = 0;
pub const RTIOW_OP_SCALE: u32 = 1;
pub const RTIOW_OP_ROTATE_Y: u32 = 2;
pub const RTIOW_OP_LINEAR_MOVE: u32 = 3;
pub const RTIOW_OP_FLIP: u32 = 4;

pub const RTIOW_MAT_LAMBERTIAN: u32 = 0;
pub const RTIOW_MAT_METAL: u32 = 1;
pub const RTIOW_MAT_DIELECTRIC: u32 = 2;
pub const RTIOW_MAT_DIFFUSE_LIGHT: u32 = 3;
pub const RTIOW_MAT_ISOTROPIC: u32 = 4;

pub const RTIOW_TEX_CONSTANT: u32 = 0;
pub const RTIOW_TEX_CHECKER: u32 = 1;
pub const RTIOW_TEX_PERLIN: u32 = 2;

pub const RTIOW_BG_BLACK: u32 = 0;
pub const RTIOW_BG_SKY_GRADIENT: u32 = 1;

pub const RTIOW_TRAVERSAL_REINDEXED: c_int = 0;
pub const RTIOW_TRAVERSAL_REFERENCE_ORDER: c_int = 1;
pub const RTIOW_TRAVERSAL_REINDEXED_EXACT: c_int = 2;

pub const RTIOW_PEER_HANDLE_BYTES: usize = 128;

#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq)]
pub struct rtiow_item_t {
    pub a: [f32; 3],
    pub a_w: u32,
    pub b: [f32; 3],
    pub b_w: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq)]
pub struct rtiow_xform_op_t {
    pub kind: u32,
    pub v: [f32; 3],
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]
pub struct rtiow_frame_t {
    pub first_op: u32,
    pub n_ops: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq)]
pub struct rtiow_material_t {
    pub kind: u32,
    pub tex: u32,
    pub albedo: [f32; 3],
    pub param: f32,
    pub reserved: [u32; 2],
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq)]
pub struct rtiow_texture_t {
    pub kind: u32,
    pub color: [f32; 3],
    pub scale: f32,
    pub child0: u32,
    pub child1: u32,
    pub reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct rtiow_scene_desc_t {
    pub abi_version: u32,
    pub n_items: u32,
    pub items: *const rtiow_item_t,
    pub n_frames: u32,
    pub n_ops: u32,
    pub frames: *const rtiow_frame_t,
    pub ops: *const rtiow_xform_op_t,
    pub n_materials: u32,
    pub n_textures: u32,
    pub materials: *const rtiow_material_t,
    pub textures: *const rtiow_texture_t,
    pub perlin_vecs: *const f32,
    pub perlin_perm: *const u8,
    pub background_kind: u32,
    pub background_c0: [f32; 3],
    pub background_c1: [f32; 3],
}

/// `src/camera.rs:6-15`, field for field (`exposure: Range<f32>` -> `time0`, `time1`).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq)]
pub struct rtiow_camera_t {
    pub origin: [f32; 3],
    pub lower_left_corner: [f32; 3],
    pub horizontal: [f32; 3],
    pub vertical: [f32; 3],
    pub u: [f32; 3],
    pub v: [f32; 3],
    pub lens_radius: f32,
    pub time0: f32,
    pub time1: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct rtiow_stats_t {
    pub trace_ms: f64,
    pub reduce_ms: f64,
    pub samples: u64,
    pub segments: u64,
    pub kernel_launches: u32,
    pub passes: u32,
    pub scene_in_smem: u32,
    pub scene_bytes: u32,
    pub grid: u32,
    pub block: u32,
    pub dyn_smem_bytes: u32,
    pub regs_per_thread: u32,
    pub accel_nodes: u32,
    pub accel_subtrees: u32,
    pub traversal: u32,
    pub kernel_profile: u32,
}

#[repr(C)]
pub struct rtiow_scene_t {
    _opaque: [u8; 0],
}
#[repr(C)]
pub struct rtiow_peer_frame_t {
    _opaque: [u8; 0],
}

extern "C" {
    pub fn rtiow_b200_abi_version() -> c_int;
    pub fn rtiow_b200_build_flavour() -> *const c_char;
    pub fn rtiow_b200_last_error() -> *const c_char;
    pub fn rtiow_b200_scene_validate(desc: *const rtiow_scene_desc_t) -> c_int;
    pub fn rtiow_b200_scene_create(desc: *const rtiow_scene_desc_t, device: c_int, out: *mut *mut rtiow_scene_t) -> c_int;
    pub fn rtiow_b200_scene_destroy(scene: *mut rtiow_scene_t);
    pub fn rtiow_b200_release_cached_memory();
    pub fn rtiow_b200_render(scene: *mut rtiow_scene_t, camera: *const rtiow_camera_t, nx: u32, ny: u32, ns: u32, seed: u64,
                             out_rgb: *mut f32) -> c_int;
    pub fn rtiow_b200_render_rows(scene: *mut rtiow_scene_t, camera: *const rtiow_camera_t, nx: u32, ny: u32, ns: u32, seed: u64,
                                  row_begin: u32, row_end: u32, out_rows: *mut f32) -> c_int;
    pub fn rtiow_b200_render_rows_device(scene: *mut rtiow_scene_t, camera: *const rtiow_camera_t, nx: u32, ny: u32, ns: u32,
                                         seed: u64, row_begin: u32, row_end: u32, d_out_rows: *mut f32,
                                         cuda_stream: *mut c_void) -> c_int;
    pub fn rtiow_b200_render_rows_strided_device(scene: *mut rtiow_scene_t, camera: *const rtiow_camera_t, nx: u32, ny: u32,
                                                 ns: u32, seed: u64, row_begin: u32, row_end: u32, row_step: u32,
                                                 band_rows: u32, d_out_rows: *mut f32, cuda_stream: *mut c_void) -> c_int;
    pub fn rtiow_b200_render_multi(scenes: *const *mut rtiow_scene_t, ngpus: c_int, camera: *const rtiow_camera_t, nx: u32,
                                   ny: u32, ns: u32, seed: u64, out_rgb: *mut f32) -> c_int;
    pub fn rtiow_b200_peer_frame_create(device: c_int, nx: u32, ny: u32, rank: u32, n_ranks: u32,
                                        out: *mut *mut rtiow_peer_frame_t) -> c_int;
    pub fn rtiow_b200_peer_frame_export(frame: *mut rtiow_peer_frame_t, handle: *mut u8) -> c_int;
    pub fn rtiow_b200_peer_frame_connect(frame: *mut rtiow_peer_frame_t, handles: *const u8) -> c_int;
    pub fn rtiow_b200_peer_frame_ptr(frame: *mut rtiow_peer_frame_t, d_frame: *mut *mut f32) -> c_int;
    pub fn rtiow_b200_peer_frame_destroy(frame: *mut rtiow_peer_frame_t);
    pub fn rtiow_b200_render_peers(scene: *mut rtiow_scene_t, camera: *const rtiow_camera_t, nx: u32, ny: u32, ns: u32,
                                        seed: u64, frame: *mut rtiow_peer_frame_t,
                                        cuda_stream: *mut c_void) -> c_int;
    pub fn rtiow_b200_render_samples(scene: *mut rtiow_scene_t, camera: *const rtiow_camera_t, nx: u32, ny: u32, ns: u32,
                                     seed: u64, row_begin: u32, row_end: u32, out_samples: *mut f32) -> c_int;
    pub fn rtiow_b200_ppm_quantise(scene: *mut rtiow_scene_t, linear: *const f32, n: usize, out: *mut u8) -> c_int;
    pub fn rtiow_b200_ppm_quantise_device(scene: *mut rtiow_scene_t, d_linear: *const f32, n: usize, d_out: *mut u8,
                                          cuda_stream: *mut c_void) -> c_int;
    pub fn rtiow_b200_render_ppm(scene: *mut rtiow_scene_t, camera: *const rtiow_camera_t, nx: u32, ny: u32, ns: u32, seed: u64,
                                 out_rgb8: *mut u8) -> c_int;
    pub fn rtiow_b200_get_stats(scene: *mut rtiow_scene_t, out: *mut rtiow_stats_t) -> c_int;
    pub fn rtiow_b200_set_tuning(scene: *mut rtiow_scene_t, cta_threads: u32, ctas_per_sm: u32, staging_mib: u32,
                                 force_global: c_int) -> c_int;
    pub fn rtiow_b200_set_specialisation(scene: *mut rtiow_scene_t, enable: c_int) -> c_int;
    pub fn rtiow_b200_set_traversal(scene: *mut rtiow_scene_t, mode: c_int) -> c_int;
}

// layout pins (same numbers as tests/test_host_cpu.py::test_abi_struct_sizes_match_header)
const _: () = assert!(std::mem::size_of::<rtiow_item_t>() == 32);
const _: () = assert!(std::mem::size_of::<rtiow_xform_op_t>() == 16);
const _: () = assert!(std::mem::size_of::<rtiow_frame_t>() == 8);
const _: () = assert!(std::mem::size_of::<rtiow_material_t>() == 32);
const _: () = assert!(std::mem::size_of::<rtiow_texture_t>() == 32);
const _: () = assert!(std::mem::size_of::<rtiow_camera_t>() == 84);
