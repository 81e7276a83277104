//! UNCOMPILED (see ../../README.md).  The flattener: the object tree of the reference
//! (`Box<dyn Object>`, `src/object.rs`) -> the plain arrays of `rtiow_scene_desc_t`.
//!
//! Line-for-line twin of `rtiow::SceneBuilder` in `rtiow-rust_b200/csrc/host/rtiow_host.cpp`, which is
//! compiled and tested against the oracle; keep the two in step.
//!
//! The stream IS the semantics: items appear in the reference's own depth-first, left-first visiting
//! order (`Bvh::hit` `src/bvh.rs:85-120`, `And::hit` `src/object.rs:396-410`, the list loop
//! `src/lib.rs:40-45`), a failed box test jumps to its `skip` link, everything else advances by one.
use std::collections::HashMap;

use rtiow_b200_sys as sys;

/// What `color()` returns when a ray escapes: `Black` is HEAD's behaviour (`src/lib.rs:100`),
/// `SkyGradient` the book-1 sky the README image was rendered with.
#[derive(Clone, Copy, Debug, PartialEq)]
pub enum Background {
    Black,
    SkyGradient { c0: [f32; 3], c1: [f32; 3] },
}

/// `src/texture.rs:6-26` as data instead of closures.
#[derive(Clone, Debug, PartialEq)]
pub enum TextureDesc {
    Constant([f32; 3]),
    Checker(Box<TextureDesc>, Box<TextureDesc>),
    Perlin { scale: f32 },
}

/// `src/material.rs:11-40` with `TextureDesc` in place of the closure.
#[derive(Clone, Debug, PartialEq)]
pub enum MaterialDesc {
    Lambertian { albedo: TextureDesc },
    Metal { albedo: [f32; 3], fuzz: f32 },
    Dielectric { ref_idx: f32 },
    DiffuseLight { emission: TextureDesc, brightness: f32 },
    Isotropic { albedo: TextureDesc },
}

/// `src/perlin.rs:5-29`: 256 gradient vectors and the three permutations, as the process generated them.
#[derive(Clone)]
pub struct PerlinTables {
    pub vecs: [[f32; 3]; 256],
    pub perm: [[u8; 256]; 3],
}

/// Implemented by every `Object` (made a supertrait of `Object` by `reference-patch/rtiow.patch`, so
/// that `Box<dyn Object>` can flatten itself): emit yourself into `b` in hit-visiting order.
pub trait Flatten {
    fn flatten(&self, b: &mut SceneBuilder);
}

/// Owner of the arrays a `rtiow_scene_desc_t` points into.
pub struct FlatScene {
    pub items: Vec<sys::rtiow_item_t>,
    pub frames: Vec<sys::rtiow_frame_t>,
    pub ops: Vec<sys::rtiow_xform_op_t>,
    pub materials: Vec<sys::rtiow_material_t>,
    pub textures: Vec<sys::rtiow_texture_t>,
    pub perlin: Option<Box<PerlinTables>>,
    pub background: Background,
}

impl FlatScene {
    /// The descriptor; valid while `self` is borrowed.
    pub fn desc(&self) -> sys::rtiow_scene_desc_t {
        let (kind, c0, c1) = match self.background {
            Background::Black => (sys::RTIOW_BG_BLACK, [0.; 3], [0.; 3]),
            Background::SkyGradient { c0, c1 } => (sys::RTIOW_BG_SKY_GRADIENT, c0, c1),
        };
        sys::rtiow_scene_desc_t {
            abi_version: sys::RTIOW_B200_ABI_VERSION,
            n_items: self.items.len() as u32,
            items: self.items.as_ptr(),
            n_frames: self.frames.len() as u32,
            n_ops: self.ops.len() as u32,
            frames: self.frames.as_ptr(),
            ops: if self.ops.is_empty() { std::ptr::null() } else { self.ops.as_ptr() },
            n_materials: self.materials.len() as u32,
            n_textures: self.textures.len() as u32,
            materials: if self.materials.is_empty() { std::ptr::null() } else { self.materials.as_ptr() },
            textures: if self.textures.is_empty() { std::ptr::null() } else { self.textures.as_ptr() },
            perlin_vecs: self.perlin.as_ref().map_or(std::ptr::null(), |t| t.vecs.as_ptr() as *const f32),
            perlin_perm: self.perlin.as_ref().map_or(std::ptr::null(), |t| t.perm.as_ptr() as *const u8),
            background_kind: kind,
            background_c0: c0,
            background_c1: c1,
        }
    }
}

/// A wrapper between the world and the current object (`Translate`, `Scale`, `RotateY`, `LinearMove`,
/// `FlipNormals`), applied outermost first to the ray and innermost first to the hit record.
type Op = sys::rtiow_xform_op_t;

fn op_key(ops: &[Op]) -> Vec<u32> {
    ops.iter().flat_map(|o| [o.kind, o.v[0].to_bits(), o.v[1].to_bits(), o.v[2].to_bits()]).collect()
}

/// Collects the arrays while objects flatten themselves.
pub struct SceneBuilder {
    items: Vec<sys::rtiow_item_t>,
    frames: Vec<sys::rtiow_frame_t>,
    ops: Vec<Op>,
    materials: Vec<sys::rtiow_material_t>,
    textures: Vec<sys::rtiow_texture_t>,
    frame_index: HashMap<Vec<u32>, u32>,
    material_index: HashMap<Vec<u32>, u32>,
    texture_index: HashMap<Vec<u32>, u32>,
    chain: Vec<Op>,           // wrappers between the world and the current object
    prefix_stack: Vec<usize>, // chain length already applied by the enclosing BBOX frame / medium
    frame_stack: Vec<u32>,
    cur_frame: u32,
    medium_depth: u32,
    medium_item: usize,
    next_medium_id: u32,
    uses_perlin: bool,
}

struct Inline {
    has_offset: bool,
    flip: bool,
    offset: [f32; 3],
    kept: usize,
}

impl Default for SceneBuilder {
    fn default() -> Self {
        Self::new()
    }
}

impl SceneBuilder {
    pub fn new() -> Self {
        let mut frame_index = HashMap::new();
        frame_index.insert(Vec::new(), 0u32);
        SceneBuilder {
            items: Vec::new(),
            frames: vec![sys::rtiow_frame_t { first_op: 0, n_ops: 0 }], // frame 0 = the world
            ops: Vec::new(),
            materials: Vec::new(),
            textures: Vec::new(),
            frame_index,
            material_index: HashMap::new(),
            texture_index: HashMap::new(),
            chain: Vec::new(),
            prefix_stack: vec![0],
            frame_stack: Vec::new(),
            cur_frame: 0,
            medium_depth: 0,
            medium_item: 0,
            next_medium_id: 0,
            uses_perlin: false,
        }
    }

    // ---- wrappers (`Translate::flatten` etc. bracket their inner object with these) ---------------
    pub fn push_op(&mut self, kind: u32, v: [f32; 3]) {
        self.chain.push(Op { kind, v });
    }
    pub fn pop_op(&mut self) {
        self.chain.pop();
    }

    fn intern_frame(&mut self, n: usize) -> u32 {
        let key = op_key(&self.chain[..n]);
        if let Some(&id) = self.frame_index.get(&key) {
            return id;
        }
        let id = self.frames.len() as u32;
        self.frames.push(sys::rtiow_frame_t { first_op: self.ops.len() as u32, n_ops: n as u32 });
        self.ops.extend_from_slice(&self.chain[..n]);
        self.frame_index.insert(key, id);
        id
    }

    fn intern_texture(&mut self, t: &TextureDesc) -> u32 {
        let mut rec = sys::rtiow_texture_t::default();
        match t {
            TextureDesc::Constant(c) => {
                rec.kind = sys::RTIOW_TEX_CONSTANT;
                rec.color = *c;
            }
            TextureDesc::Checker(t0, t1) => {
                rec.kind = sys::RTIOW_TEX_CHECKER; // children first: child index < parent index
                rec.child0 = self.intern_texture(t0);
                rec.child1 = self.intern_texture(t1);
            }
            TextureDesc::Perlin { scale } => {
                rec.kind = sys::RTIOW_TEX_PERLIN;
                rec.scale = *scale;
                self.uses_perlin = true;
            }
        }
        let key = vec![rec.kind, rec.color[0].to_bits(), rec.color[1].to_bits(), rec.color[2].to_bits(), rec.scale.to_bits(),
                       rec.child0, rec.child1];
        if let Some(&id) = self.texture_index.get(&key) {
            return id;
        }
        let id = self.textures.len() as u32;
        self.textures.push(rec);
        self.texture_index.insert(key, id);
        id
    }

    fn intern_material(&mut self, m: &MaterialDesc) -> u32 {
        let mut rec = sys::rtiow_material_t::default();
        match m {
            MaterialDesc::Lambertian { albedo } => {
                rec.kind = sys::RTIOW_MAT_LAMBERTIAN;
                rec.tex = self.intern_texture(albedo);
            }
            MaterialDesc::Metal { albedo, fuzz } => {
                rec.kind = sys::RTIOW_MAT_METAL;
                rec.albedo = *albedo;
                rec.param = *fuzz;
            }
            MaterialDesc::Dielectric { ref_idx } => {
                rec.kind = sys::RTIOW_MAT_DIELECTRIC;
                rec.param = *ref_idx;
            }
            MaterialDesc::DiffuseLight { emission, brightness } => {
                rec.kind = sys::RTIOW_MAT_DIFFUSE_LIGHT;
                rec.tex = self.intern_texture(emission);
                rec.param = *brightness;
            }
            MaterialDesc::Isotropic { albedo } => {
                rec.kind = sys::RTIOW_MAT_ISOTROPIC;
                rec.tex = self.intern_texture(albedo);
            }
        }
        let key = vec![rec.kind, rec.tex, rec.albedo[0].to_bits(), rec.albedo[1].to_bits(), rec.albedo[2].to_bits(), rec.param.to_bits()];
        if let Some(&id) = self.material_index.get(&key) {
            return id;
        }
        let id = self.materials.len() as u32;
        self.materials.push(rec);
        self.material_index.insert(key, id);
        id
    }

    /// Folds the innermost `FlipNormals` / `Translate` wrappers (those not already applied by the
    /// enclosing BBOX frame or medium) into the primitive's own record.  `FlipNormals` only negates the
    /// normal and `Translate` only shifts origin and `p`, so they commute and the arithmetic is unchanged.
    fn split_inline(&self, allow_offset: bool) -> Inline {
        let mut r = Inline { has_offset: false, flip: false, offset: [0.; 3], kept: self.chain.len() };
        let prefix = *self.prefix_stack.last().unwrap();
        while r.kept > prefix {
            let op = self.chain[r.kept - 1];
            if op.kind == sys::RTIOW_OP_FLIP {
                r.flip = !r.flip;
            } else if op.kind == sys::RTIOW_OP_TRANSLATE && allow_offset && !r.has_offset {
                r.has_offset = true;
                r.offset = op.v;
            } else {
                break;
            }
            r.kept -= 1;
        }
        r
    }

    // ---- primitives -------------------------------------------------------------------------------
    /// `Sphere` (`src/object.rs:74-119`).
    pub fn emit_sphere(&mut self, radius: f32, m: &MaterialDesc) {
        let inl = self.split_inline(true);
        let frame = self.intern_frame(inl.kept);
        let flags = if inl.has_offset { sys::RTIOW_FLAG_HAS_OFFSET } else { 0 } | if inl.flip { sys::RTIOW_FLAG_FLIP } else { 0 };
        let mat = self.intern_material(m);
        self.items.push(sys::rtiow_item_t {
            a: [radius, 0., 0.],
            a_w: sys::RTIOW_ITEM_SPHERE | (frame << 4),
            b: inl.offset,
            b_w: mat | (flags << 24),
        });
    }

    /// `Rect<A>` (`src/object.rs:131-234`); `axis` = 0 X, 1 Y, 2 Z (`A::AXIS`).
    pub fn emit_rect(&mut self, axis: u32, range0: (f32, f32), range1: (f32, f32), k: f32, m: &MaterialDesc) {
        let inl = self.split_inline(false);
        let frame = self.intern_frame(inl.kept);
        let flags = if inl.flip { sys::RTIOW_FLAG_FLIP } else { 0 } | (axis << sys::RTIOW_FLAG_AXIS_SHIFT);
        let mat = self.intern_material(m);
        self.items.push(sys::rtiow_item_t {
            a: [k, range0.0, range0.1],
            a_w: sys::RTIOW_ITEM_RECT | (frame << 4),
            b: [range1.0, range1.1, 0.],
            b_w: mat | (flags << 24),
        });
    }

    // ---- Bvh (`src/bvh.rs`) -----------------------------------------------------------------------
    /// A node's box; returns a token for `end_bbox`.
    pub fn begin_bbox(&mut self, min: [f32; 3], max: [f32; 3]) -> usize {
        self.items.push(sys::rtiow_item_t { a: min, a_w: sys::RTIOW_ITEM_BBOX, b: max, b_w: 0 });
        self.items.len() - 1
    }
    pub fn end_bbox(&mut self, token: usize) {
        self.items[token].a_w = sys::RTIOW_ITEM_BBOX | ((self.items.len() as u32) << 4); // skip link
    }
    /// A `Bvh` root: its boxes live in the frame of the wrappers around it.
    pub fn begin_subtree(&mut self) -> Result<(), String> {
        if self.medium_depth > 0 {
            if self.chain.len() != *self.prefix_stack.last().unwrap() {
                return Err("a Bvh used as ConstantMedium boundary must not be wrapped inside the medium; wrap the ConstantMedium instead".into());
            }
            self.frame_stack.push(self.cur_frame);
            self.prefix_stack.push(self.chain.len());
            return Ok(());
        }
        self.frame_stack.push(self.cur_frame);
        if self.chain.len() != *self.prefix_stack.last().unwrap() {
            let f = self.intern_frame(self.chain.len());
            self.items.push(sys::rtiow_item_t { a_w: sys::RTIOW_ITEM_SET_FRAME | (f << 4), ..Default::default() });
            self.cur_frame = f;
        }
        self.prefix_stack.push(self.chain.len());
        Ok(())
    }
    pub fn end_subtree(&mut self) {
        self.prefix_stack.pop();
        let prev = self.frame_stack.pop().unwrap();
        if prev != self.cur_frame {
            self.items.push(sys::rtiow_item_t { a_w: sys::RTIOW_ITEM_SET_FRAME | (prev << 4), ..Default::default() });
            self.cur_frame = prev;
        }
    }

    // ---- ConstantMedium<O> (`src/object.rs:533-580`) ----------------------------------------------
    /// The medium item; the boundary object then flattens itself (any primitives, or a `Bvh`), and
    /// `end_medium` closes the run.  Media are numbered in flattening order: that number selects the
    /// RNG stream of the medium's free-path draw (DESIGN.md "RNG contract").
    pub fn begin_medium(&mut self, density: f32, m: &MaterialDesc) -> Result<(), String> {
        if self.medium_depth > 0 {
            return Err("a ConstantMedium inside a ConstantMedium boundary is not supported".into());
        }
        let frame = self.intern_frame(self.chain.len());
        let mat = self.intern_material(m);
        let id = self.next_medium_id;
        self.next_medium_id += 1;
        self.items.push(sys::rtiow_item_t {
            a: [density, f32::from_bits(id), 0.],
            a_w: sys::RTIOW_ITEM_MEDIUM | (frame << 4),
            b: [0.; 3],
            b_w: mat,
        });
        self.medium_item = self.items.len() - 1;
        self.prefix_stack.push(self.chain.len());
        self.medium_depth += 1;
        Ok(())
    }
    pub fn end_medium(&mut self) -> Result<(), String> {
        self.medium_depth -= 1;
        self.prefix_stack.pop();
        if self.items.len() == self.medium_item + 1 {
            return Err("ConstantMedium boundary flattened to nothing".into());
        }
        self.items[self.medium_item].a[2] = f32::from_bits(self.items.len() as u32); // index after the boundary run
        Ok(())
    }

    /// Appends END and hands the arrays over.
    pub fn finish(mut self, perlin: Option<&PerlinTables>, background: Background) -> Result<FlatScene, String> {
        self.items.push(sys::rtiow_item_t::default()); // RTIOW_ITEM_END
        if self.uses_perlin && perlin.is_none() {
            return Err("Perlin texture without Perlin tables".into());
        }
        Ok(FlatScene {
            items: self.items,
            frames: self.frames,
            ops: self.ops,
            materials: self.materials,
            textures: self.textures,
            perlin: perlin.map(|t| Box::new(t.clone())),
            background,
        })
    }
}

/// `impl World for [Box<dyn Object>]` (`src/lib.rs:33-49`): the list in order.  For a top-level
/// `Bvh` (`src/lib.rs:51-55`) pass a one-element slice holding it.
pub fn flatten_world<O: Flatten>(world: &[O], perlin: Option<&PerlinTables>, background: Background) -> Result<FlatScene, String> {
    let mut b = SceneBuilder::new();
    for o in world {
        o.flatten(&mut b);
    }
    b.finish(perlin, background)
}
