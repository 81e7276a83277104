//! UNCOMPILED (see ../../README.md).  Safe wrappers over the C ABI: a scene resident on a B200 and
//! `par_cast` on it.  Everything `unsafe` of the Rust side lives in this file.
use std::ffi::CStr;
use std::fmt;

use rtiow_b200_sys as sys;

use crate::flat::FlatScene;

/// An `RTIOW_ERR_*` code with the library's message (`rtiow_b200_last_error`).
#[derive(Debug, Clone)]
pub struct Error {
    pub code: i32,
    pub message: String,
}

impl fmt::Display for Error {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        write!(f, "[rtiow_b200 error {}] {}", self.code, self.message)
    }
}
impl std::error::Error for Error {}

fn check(rc: i32) -> Result<(), Error> {
    if rc == sys::RTIOW_OK {
        return Ok(());
    }
    // SAFETY: the library returns a pointer to a thread-local NUL-terminated string that lives until the next call.
    let message = unsafe { CStr::from_ptr(sys::rtiow_b200_last_error()) }.to_string_lossy().into_owned();
    Err(Error { code: rc, message })
}

/// How `Bvh` subtrees are walked; every mode gives the same image (`RTIOW_TRAVERSAL_*`).
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Traversal {
    Reindexed = 0,
    ReferenceOrder = 1,
    ReindexedExact = 2,
}

/// A scene uploaded to one device (`rtiow_b200_scene_create` .. `rtiow_b200_scene_destroy`).
/// One render at a time per scene (the C handle owns one set of device work buffers), hence `&mut self` on renders
/// and no `Sync`.
pub struct GpuScene {
    raw: *mut sys::rtiow_scene_t,
}

// SAFETY: the handle may move between threads; the library only requires that one thread uses it at a time.
unsafe impl Send for GpuScene {}

impl GpuScene {
    /// Validates and uploads `flat` to CUDA device `device`.  Fails with `RTIOW_ERR_NO_DEVICE` without an sm_100 GPU:
    /// there is deliberately no CPU fallback behind this call (the crate's own `par_cast` is the CPU path).
    pub fn new(flat: &FlatScene, device: i32) -> Result<GpuScene, Error> {
        let desc = flat.desc();
        let mut raw = std::ptr::null_mut();
        // SAFETY: `desc` points into `flat`, which outlives the call; the library copies everything it keeps.
        check(unsafe { sys::rtiow_b200_scene_create(&desc, device, &mut raw) })?;
        Ok(GpuScene { raw })
    }

    pub fn set_traversal(&mut self, mode: Traversal) -> Result<(), Error> {
        // SAFETY: `raw` is a live handle.
        check(unsafe { sys::rtiow_b200_set_traversal(self.raw, mode as i32) })
    }

    pub fn stats(&mut self) -> Result<sys::rtiow_stats_t, Error> {
        let mut st = sys::rtiow_stats_t::default();
        // SAFETY: `raw` is a live handle, `st` a valid out-pointer.
        check(unsafe { sys::rtiow_b200_get_stats(self.raw, &mut st) })?;
        Ok(st)
    }

    pub(crate) fn raw(&self) -> *mut sys::rtiow_scene_t {
        self.raw
    }
}

impl Drop for GpuScene {
    fn drop(&mut self) {
        // SAFETY: `raw` came from rtiow_b200_scene_create and is destroyed exactly once.
        unsafe { sys::rtiow_b200_scene_destroy(self.raw) }
    }
}

/// `par_cast(nx, ny, ns, &camera, world)` (`src/lib.rs:363-376`) on the GPU, with the explicit seed the reference's
/// `cast` takes as `&mut impl Rng` (`src/lib.rs:378-397`).  Returns `Image` flattened: `ny * nx * 3` floats, row 0 =
/// top scanline, linear, already divided by `ns` — wrap as `Image(rows)` with `chunks(nx * 3)`.
pub fn gpu_cast(nx: usize, ny: usize, ns: usize, camera: &sys::rtiow_camera_t, scene: &mut GpuScene, seed: u64) -> Result<Vec<f32>, Error> {
    let mut out = vec![0f32; nx * ny * 3];
    // SAFETY: `out` has exactly the ny*nx*3 floats the call writes; `camera` is a valid #[repr(C)] record.
    check(unsafe { sys::rtiow_b200_render(scene.raw(), camera, nx as u32, ny as u32, ns as u32, seed, out.as_mut_ptr()) })?;
    Ok(out)
}

/// The same over several GPUs from this one thread: `scenes[g]` = the same `FlatScene` uploaded to device `g`.
/// Bands of scanlines are dealt round-robin; every GPU's sample fold stores its rows straight into GPU 0's frame over
/// NVLink (`rtiow_b200_render_multi`).  Bit-identical to `gpu_cast`.
pub fn gpu_cast_multi(nx: usize, ny: usize, ns: usize, camera: &sys::rtiow_camera_t, scenes: &mut [GpuScene], seed: u64) -> Result<Vec<f32>, Error> {
    let handles: Vec<*mut sys::rtiow_scene_t> = scenes.iter().map(|s| s.raw()).collect();
    let mut out = vec![0f32; nx * ny * 3];
    // SAFETY: `handles` holds `scenes.len()` live handles; `out` has the ny*nx*3 floats the call writes.
    check(unsafe {
        sys::rtiow_b200_render_multi(handles.as_ptr(), handles.len() as i32, camera, nx as u32, ny as u32, ns as u32, seed, out.as_mut_ptr())
    })?;
    Ok(out)
}

/// `par_cast` + `print_ppm`'s `sqrt` / `to_u8` (`src/lib.rs:344-361`) without the float frame leaving the device:
/// `ny * nx * 3` bytes, the numbers `print_ppm` writes.
pub fn gpu_cast_ppm(nx: usize, ny: usize, ns: usize, camera: &sys::rtiow_camera_t, scene: &mut GpuScene, seed: u64) -> Result<Vec<u8>, Error> {
    let mut out = vec![0u8; nx * ny * 3];
    // SAFETY: `out` has exactly the ny*nx*3 bytes the call writes.
    check(unsafe { sys::rtiow_b200_render_ppm(scene.raw(), camera, nx as u32, ny as u32, ns as u32, seed, out.as_mut_ptr()) })?;
    Ok(out)
}
