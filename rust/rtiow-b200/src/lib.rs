//! UNCOMPILED (see ../../README.md).
//!
//! The B200 drop-in for `rtiow::par_cast` (reference `src/lib.rs:363-376`):
//!
//! ```ignore
//! // before:  let image = rtiow::par_cast(NX, NY, NS, &camera, world.as_slice());
//! let scene = rtiow_b200::GpuScene::new(&rtiow_b200::flat::flatten_world(&world, &tables, Background::Black), 0)?;
//! let image = rtiow_b200::gpu_cast(NX, NY, NS, &camera.to_repr_c(), &scene, 0xDEADBEEF)?;   // Vec<f32>, ny*nx*3, row 0 = top
//! ```
//!
//! This crate does not depend on the reference crate: the reference depends on THIS crate for the
//! `Flatten` trait (a supertrait of its `Object`, see `reference-patch/`), implements it for its own
//! types, and stays `#![forbid(unsafe_code)]`; all `unsafe` is in `gpu.rs` here and in `rtiow-b200-sys`.
pub mod flat;
pub mod gpu;

pub use flat::{Background, FlatScene, Flatten, MaterialDesc, PerlinTables, SceneBuilder, TextureDesc};
pub use gpu::{gpu_cast, gpu_cast_multi, gpu_cast_ppm, Error, GpuScene, Traversal};
pub use rtiow_b200_sys::rtiow_camera_t as CameraReprC;
