//! UNCOMPILED (see ../README.md).  Drops into the reference as `src/flat.rs` (`mod flat;` in `src/lib.rs`):
//! `impl Flatten for` every `Object` implementor, so that a scene built with the crate's own types can be handed to
//! the B200 library.  Each impl mirrors the member of the same name in `rtiow-rust_b200/csrc/host/rtiow_host.cpp`
//! (`rtiow::object::*::flatten`), which is compiled and tested against the oracle.
//!
//! Needs the edits of `rtiow.patch`: `Object: Flatten`, the `Texture` enum (closures cannot be read back,
//! `src/texture.rs:6`), and accessors for the private fields used below.
use rtiow_b200::flat::{Flatten, MaterialDesc, SceneBuilder, TextureDesc};
use rtiow_b200::CameraReprC;
use rtiow_b200_sys as sys;

use crate::bvh::{Bvh, BvhContents};
use crate::camera::Camera;
use crate::material::Material;
use crate::object::*;
use crate::texture::Texture;
use crate::vec3::{Axis, Vec3};

fn v3(v: Vec3) -> [f32; 3] {
    [v.0, v.1, v.2]
}

fn axis_index(a: Axis) -> u32 {
    match a {
        Axis::X => 0,
        Axis::Y => 1,
        Axis::Z => 2,
    }
}

/// `Texture` as data (after `rtiow.patch`: `pub enum Texture { Constant(Vec3), Checker(Box<Texture>, Box<Texture>),
/// Perlin { scale: f32 } }` with `fn eval(&self, p: Vec3) -> Vec3` doing what the closures did, `src/texture.rs:8-26`).
impl From<&Texture> for TextureDesc {
    fn from(t: &Texture) -> TextureDesc {
        match t {
            Texture::Constant(c) => TextureDesc::Constant(v3(*c)),
            Texture::Checker(t0, t1) => TextureDesc::Checker(Box::new((&**t0).into()), Box::new((&**t1).into())),
            Texture::Perlin { scale } => TextureDesc::Perlin { scale: *scale },
        }
    }
}

/// `src/material.rs:11-40`.
impl From<&Material> for MaterialDesc {
    fn from(m: &Material) -> MaterialDesc {
        match m {
            Material::Lambertian { albedo } => MaterialDesc::Lambertian { albedo: albedo.into() },
            Material::Metal { albedo, fuzz } => MaterialDesc::Metal { albedo: v3(*albedo), fuzz: *fuzz },
            Material::Dielectric { ref_idx } => MaterialDesc::Dielectric { ref_idx: *ref_idx },
            Material::DiffuseLight { emission, brightness } => MaterialDesc::DiffuseLight { emission: emission.into(), brightness: *brightness },
            Material::Isotropic { albedo } => MaterialDesc::Isotropic { albedo: albedo.into() },
        }
    }
}

// 1. Box<dyn Object> (src/object.rs:42-54): forwards, like its `hit`.
impl Flatten for Box<dyn Object> {
    fn flatten(&self, b: &mut SceneBuilder) {
        (**self).flatten(b)
    }
}

// 2. Sphere (src/object.rs:74-119)
impl Flatten for Sphere {
    fn flatten(&self, b: &mut SceneBuilder) {
        b.emit_sphere(self.radius, &(&self.material).into());
    }
}

// 3. Rect<A> (src/object.rs:131-234): "other two" axes are alphabetical, which is also the item's layout.
impl<A: StaticAxis> Flatten for Rect<A> {
    fn flatten(&self, b: &mut SceneBuilder) {
        b.emit_rect(axis_index(A::AXIS), (self.range0.start, self.range0.end), (self.range1.start, self.range1.end), self.k,
                    &(&self.material).into());
    }
}

// 4. FlipNormals<O> (src/object.rs:238-258)
impl<O: Object> Flatten for FlipNormals<O> {
    fn flatten(&self, b: &mut SceneBuilder) {
        b.push_op(sys::RTIOW_OP_FLIP, [0.; 3]);
        self.0.flatten(b);
        b.pop_op();
    }
}

// 5. Translate<O> (src/object.rs:261-292)
impl<O: Object> Flatten for Translate<O> {
    fn flatten(&self, b: &mut SceneBuilder) {
        b.push_op(sys::RTIOW_OP_TRANSLATE, v3(self.offset));
        self.object.flatten(b);
        b.pop_op();
    }
}

// 6. Scale<O> (src/object.rs:295-328)
impl<O: Object> Flatten for Scale<O> {
    fn flatten(&self, b: &mut SceneBuilder) {
        b.push_op(sys::RTIOW_OP_SCALE, v3(self.factor));
        self.object.flatten(b);
        b.pop_op();
    }
}

// 7. RotateY<O> (src/object.rs:335-390); `sin_cos()` is the accessor the patch adds for the two private fields.
impl<O: Object> Flatten for RotateY<O> {
    fn flatten(&self, b: &mut SceneBuilder) {
        let (sin_theta, cos_theta) = self.sin_cos();
        b.push_op(sys::RTIOW_OP_ROTATE_Y, [sin_theta, cos_theta, 0.]);
        self.object.flatten(b);
        b.pop_op();
    }
}

// 8. And<T, S> (src/object.rs:394-417): hit order is .0 then .1 with the tightened range, i.e. stream order.
//    rect_prism (object.rs:420-473) is And-of-And of six rects: the library fuses such a run into one prism record.
impl<T: Object, S: Object> Flatten for And<T, S> {
    fn flatten(&self, b: &mut SceneBuilder) {
        self.0.flatten(b);
        self.1.flatten(b);
    }
}

// 9. LinearMove<O> (src/object.rs:489-528)
impl<O: Object> Flatten for LinearMove<O> {
    fn flatten(&self, b: &mut SceneBuilder) {
        b.push_op(sys::RTIOW_OP_LINEAR_MOVE, v3(self.motion));
        self.object.flatten(b);
        b.pop_op();
    }
}

// 10. ConstantMedium<O> (src/object.rs:533-580): the boundary is any Object; it flattens itself between
//     begin_medium and end_medium and becomes the medium's boundary run.
impl<O: Object> Flatten for ConstantMedium<O> {
    fn flatten(&self, b: &mut SceneBuilder) {
        b.begin_medium(self.density, &(&self.material).into()).expect("ConstantMedium nesting the device format cannot express");
        self.boundary.flatten(b);
        b.end_medium().expect("ConstantMedium with an empty boundary");
    }
}

// 11. Bvh (src/bvh.rs): every node's box with a skip link, children left first (Bvh::hit, bvh.rs:85-120).
//     `bounding_box_ref()` / `contents()` are the accessors the patch adds for the private fields.
impl Bvh {
    fn flatten_node(&self, b: &mut SceneBuilder) {
        let bb = self.bounding_box_ref();
        let token = b.begin_bbox(v3(bb.min), v3(bb.max));
        match self.contents() {
            BvhContents::Leaf(obj) => obj.flatten(b),
            BvhContents::Node { left, right } => {
                left.flatten_node(b);
                right.flatten_node(b);
            }
        }
        b.end_bbox(token);
    }
}

impl Flatten for Bvh {
    fn flatten(&self, b: &mut SceneBuilder) {
        b.begin_subtree().expect("Bvh wrapped inside a ConstantMedium boundary");
        self.flatten_node(b);
        b.end_subtree();
    }
}

/// `Camera` field for field (`src/camera.rs:6-15`): what `rtiow_camera_t` is.
impl Camera {
    pub fn to_repr_c(&self) -> CameraReprC {
        CameraReprC {
            origin: v3(self.origin),
            lower_left_corner: v3(self.lower_left_corner),
            horizontal: v3(self.horizontal),
            vertical: v3(self.vertical),
            u: v3(self.u),
            v: v3(self.v),
            lens_radius: self.lens_radius,
            time0: self.exposure.start,
            time1: self.exposure.end,
        }
    }
}
