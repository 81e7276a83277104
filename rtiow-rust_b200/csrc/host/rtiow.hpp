// Host-side mirror of the reference crate's public surface for the render path, in C++
// (the reference is Rust; no Rust toolchain exists in this environment).
//
// Names, argument meaning and error behaviour follow cbiffle/rtiow-rust:
//   rtiow::Vec3                         src/vec3.rs
//   rtiow::camera::Camera::look         src/camera.rs:18-50
//   rtiow::material::Material           src/material.rs:11-40
//   rtiow::texture::{constant,checker,perlin}   src/texture.rs:8-26
//   rtiow::object::{Sphere,Rect,FlipNormals,Translate,Scale,RotateY,And,LinearMove,ConstantMedium,
//                   rect_prism,rotate_y}        src/object.rs
//   rtiow::bvh::{Bvh,from_scene}        src/bvh.rs
//   rtiow::{World,Image,par_cast,cast,print_ppm,cornell_box,cornell_box_with_boxes}   src/lib.rs
//
// What differs: objects do not carry `hit` (ray intersection runs on the device); instead every
// object can `flatten` itself into the C ABI's traversal stream (include/rtiow_b200.h), which is
// the introspection hook the Rust crate would need as well (its `dyn Object` / closure textures
// are opaque, src/object.rs:15-40, src/texture.rs:6).
#pragma once
#include <array>
#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/rtiow_b200.h"

namespace rtiow {

struct Vec3 {
    float x = 0.f, y = 0.f, z = 0.f;
    Vec3() = default;
    Vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    static Vec3 from(float v) { return Vec3(v, v, v); }
    float operator[](int axis) const { return axis == 0 ? x : (axis == 1 ? y : z); }
    float& operator[](int axis) { return axis == 0 ? x : (axis == 1 ? y : z); }
    float dot(Vec3 o) const;
    Vec3 cross(Vec3 o) const;
    float length() const;
    Vec3 into_unit() const;
};
Vec3 operator+(Vec3 a, Vec3 b);
Vec3 operator-(Vec3 a, Vec3 b);
Vec3 operator*(Vec3 a, Vec3 b);
Vec3 operator/(Vec3 a, Vec3 b);
Vec3 operator*(float s, Vec3 v);
Vec3 operator+(float s, Vec3 v);
Vec3 operator/(Vec3 a, float s);
Vec3 operator-(Vec3 a);

struct Range {  // std::ops::Range<f32>
    float start = 0.f, end = 0.f;
};

struct Aabb {  // src/aabb.rs (hit() lives on the device)
    Vec3 min, max;
    Aabb merge(const Aabb& other) const;
    std::array<Vec3, 8> corners() const;
};

// ------------------------------------------------------------------------------------------ rng
// rand 0.6.5's SmallRng (rand_pcg::Pcg64Mcg) as used by src/main.rs:333 for scene generation.
class SmallRng {
 public:
    static SmallRng seed_from_u64(uint64_t seed);
    uint32_t next_u32();
    uint64_t next_u64();
    float gen_f32();                       // rng.gen::<f32>()
    Vec3 gen_vec3();                       // rng.gen::<Vec3>()   src/vec3.rs:209-214
    float gen_range(float low, float high);
    size_t gen_range(size_t low, size_t high);

 private:
    unsigned __int128 state_ = 1;
};

// ------------------------------------------------------------------------------------- textures
namespace texture {
struct Node;
using Texture = std::shared_ptr<const Node>;  // the crate's `Arc<dyn Fn(Vec3) -> Vec3>` as data
struct Node {
    uint32_t kind;  // RTIOW_TEX_*
    Vec3 color;
    float scale = 0.f;
    Texture t0, t1;
};
Texture constant(Vec3 color);
Texture checker(Texture t0, Texture t1);
Texture perlin(float scale);
}  // namespace texture

struct PerlinTables {  // src/perlin.rs:5-29 (seeded instead of thread_rng())
    float vecs[256][3];
    uint8_t perm[3][256];
    static std::shared_ptr<const PerlinTables> generate(uint64_t scene_seed);
};

// ------------------------------------------------------------------------------------ materials
namespace material {
struct Material {
    uint32_t kind = RTIOW_MAT_LAMBERTIAN;
    texture::Texture tex;
    Vec3 albedo;
    float param = 0.f;
    static Material Lambertian(texture::Texture albedo);
    static Material Metal(Vec3 albedo, float fuzz);
    static Material Dielectric(float ref_idx);
    static Material DiffuseLight(texture::Texture emission, float brightness);
    static Material Isotropic(texture::Texture albedo);
};
}  // namespace material

// ------------------------------------------------------------------------------------- flatten
// Collects the C ABI's arrays while objects flatten themselves in hit-visiting order.
class SceneBuilder {
 public:
    SceneBuilder();
    // called by Object::flatten implementations
    void push_op(uint32_t kind, Vec3 v);
    void pop_op();
    void emit_sphere(float radius, const material::Material& m);
    void emit_rect(int axis, Range r0, Range r1, float k, const material::Material& m);
    size_t begin_bbox(const Aabb& box);  // returns a token for end_bbox
    void end_bbox(size_t token);
    void begin_subtree();                // a Bvh root: switches the BBOX frame if wrappers intervened
    void end_subtree();
    void begin_medium(float density, const material::Material& m, uint32_t medium_id);
    void end_medium();

    void set_background(uint32_t kind, Vec3 c0, Vec3 c1);
    void set_perlin(std::shared_ptr<const PerlinTables> t) { perlin_ = std::move(t); }
    // Appends END and returns a descriptor pointing into this builder (valid while it lives).
    const rtiow_scene_desc_t& finish();

    const std::vector<rtiow_item_t>& items() const { return items_; }

 private:
    uint32_t intern_frame(size_t n_ops_of_chain);
    uint32_t intern_material(const material::Material& m);
    uint32_t intern_texture(const texture::Texture& t);
    struct Inline { bool has_offset = false; bool flip = false; Vec3 offset; size_t kept = 0; };
    Inline split_inline(bool allow_offset) const;

    std::vector<rtiow_item_t> items_;
    std::vector<rtiow_frame_t> frames_;
    std::vector<rtiow_xform_op_t> ops_;
    std::vector<rtiow_material_t> materials_;
    std::vector<rtiow_texture_t> textures_;
    std::map<std::string, uint32_t> frame_index_, material_index_, texture_index_;
    std::vector<rtiow_xform_op_t> chain_;  // wrappers between the world and the current object
    std::vector<size_t> prefix_stack_;     // chain length already applied by the enclosing BBOX frame / medium
    std::vector<uint32_t> frame_stack_;
    uint32_t cur_frame_ = 0;
    int medium_depth_ = 0;
    size_t medium_item_ = 0;
    std::shared_ptr<const PerlinTables> perlin_;
    rtiow_scene_desc_t desc_{};
    bool finished_ = false;
};

// -------------------------------------------------------------------------------------- objects
namespace object {

class Object {  // src/object.rs:15-40
 public:
    virtual ~Object() = default;
    virtual Aabb bounding_box(Range exposure) const = 0;
    virtual void flatten(SceneBuilder& b) const = 0;
};
using Box = std::unique_ptr<Object>;  // Box<dyn Object>

enum class StaticAxis { X = 0, Y = 1, Z = 2 };
constexpr StaticAxis StaticX = StaticAxis::X, StaticY = StaticAxis::Y, StaticZ = StaticAxis::Z;

struct Sphere : Object {  // object.rs:74-119
    float radius;
    material::Material material;
    Sphere(float r, material::Material m) : radius(r), material(std::move(m)) {}
    Aabb bounding_box(Range) const override;
    void flatten(SceneBuilder& b) const override;
};

struct Rect : Object {  // object.rs:131-234
    StaticAxis orthogonal_to;
    Range range0, range1;
    float k;
    material::Material material;
    Rect(StaticAxis a, Range r0, Range r1, float k_, material::Material m)
        : orthogonal_to(a), range0(r0), range1(r1), k(k_), material(std::move(m)) {}
    Aabb bounding_box(Range) const override;
    void flatten(SceneBuilder& b) const override;
};

struct FlipNormals : Object {  // object.rs:238-258
    Box inner;
    explicit FlipNormals(Box o) : inner(std::move(o)) {}
    Aabb bounding_box(Range e) const override;
    void flatten(SceneBuilder& b) const override;
};

struct Translate : Object {  // object.rs:261-292
    Vec3 offset;
    Box object;
    Translate(Vec3 off, Box o) : offset(off), object(std::move(o)) {}
    Aabb bounding_box(Range e) const override;
    void flatten(SceneBuilder& b) const override;
};

struct Scale : Object {  // object.rs:295-328
    Vec3 factor;
    Box object;
    Scale(Vec3 f, Box o) : factor(f), object(std::move(o)) {}
    Aabb bounding_box(Range e) const override;
    void flatten(SceneBuilder& b) const override;
};

struct RotateY : Object {  // object.rs:335-390; build with rotate_y()
    Box object;
    float sin_theta, cos_theta;
    RotateY(Box o, float s, float c) : object(std::move(o)), sin_theta(s), cos_theta(c) {}
    Aabb bounding_box(Range e) const override;
    void flatten(SceneBuilder& b) const override;
};
Box rotate_y(float degrees, Box object);  // object.rs:477-484

struct And : Object {  // object.rs:394-417
    Box first, second;
    And(Box a, Box b) : first(std::move(a)), second(std::move(b)) {}
    Aabb bounding_box(Range e) const override;
    void flatten(SceneBuilder& b) const override;
};
Box rect_prism(Vec3 p0, Vec3 p1, const material::Material& material);  // object.rs:420-473

struct LinearMove : Object {  // object.rs:489-528
    Box object;
    Vec3 motion;
    LinearMove(Box o, Vec3 m) : object(std::move(o)), motion(m) {}
    Aabb bounding_box(Range e) const override;
    void flatten(SceneBuilder& b) const override;
};

struct ConstantMedium : Object {  // object.rs:533-580
    Box boundary;
    float density;
    material::Material material;
    // Which RNG stream the medium's free-path draw uses (DESIGN.md "RNG contract"): media are
    // numbered in construction order within a scene.
    uint32_t medium_id;
    ConstantMedium(Box b, float d, material::Material m, uint32_t id)
        : boundary(std::move(b)), density(d), material(std::move(m)), medium_id(id) {}
    Aabb bounding_box(Range e) const override;
    void flatten(SceneBuilder& b) const override;
};

}  // namespace object

// ------------------------------------------------------------------------------------------ bvh
namespace bvh {
class Bvh : public object::Object {  // src/bvh.rs
 public:
    // Bvh::new (bvh.rs:22-81).  Throws std::runtime_error("Can't create a BVH from zero objects.")
    // like the reference panics (bvh.rs:60).  Equal sort keys keep their input order.
    Bvh(std::vector<object::Box> objs, Range exposure);
    Aabb bounding_box(Range) const override { return bounding_box_; }
    void flatten(SceneBuilder& b) const override;
    size_t size() const { return size_; }
    size_t node_count() const;

 private:
    void flatten_node(SceneBuilder& b) const;
    Aabb bounding_box_;
    size_t size_ = 0;
    std::unique_ptr<Bvh> left_, right_;
    object::Box leaf_;
};
std::unique_ptr<Bvh> from_scene(std::vector<object::Box> scene, Range exposure);  // bvh.rs:128
}  // namespace bvh

// --------------------------------------------------------------------------------------- camera
namespace camera {
struct Camera {  // src/camera.rs:6-15
    Vec3 origin, lower_left_corner, horizontal, vertical, u, v;
    float lens_radius = 0.f;
    Range exposure;
    static Camera look(Vec3 look_from, Vec3 look_at, Vec3 up, float fov, float aspect, float aperture, float focus_dist,
                       Range exposure);  // camera.rs:18-50
    rtiow_camera_t to_repr_c() const;
};
}  // namespace camera

// ------------------------------------------------------------------------------ world + casting
enum class Background { Black = RTIOW_BG_BLACK, SkyGradient = RTIOW_BG_SKY_GRADIENT };

// `impl World for [Box<dyn Object>]` (lib.rs:33-49) and `impl World for bvh::Bvh` (lib.rs:51-55).
class World {
 public:
    explicit World(std::vector<object::Box> list) : list_(std::move(list)) {}
    explicit World(std::unique_ptr<bvh::Bvh> bvh) : bvh_(std::move(bvh)) {}
    void flatten(SceneBuilder& b) const;
    bool is_bvh() const { return bvh_ != nullptr; }
    size_t len() const { return bvh_ ? bvh_->size() : list_.size(); }

 private:
    std::vector<object::Box> list_;
    std::unique_ptr<bvh::Bvh> bvh_;
};

struct Image {  // lib.rs:321: rows top first
    size_t nx = 0, ny = 0;
    std::vector<float> rgb;  // ny * nx * 3
};

struct CastOptions {
    uint64_t seed = 0xDEADBEEFull;
    Background background = Background::Black;
    std::shared_ptr<const PerlinTables> perlin;  // required if any texture::perlin is reachable
    int device = 0;
};

// par_cast (lib.rs:363-376) on the GPU through the C ABI.  Throws std::runtime_error carrying
// rtiow_b200_last_error() on failure; there is no CPU path.
Image par_cast(size_t nx, size_t ny, size_t ns, const camera::Camera& camera, const World& world,
               const CastOptions& opts = CastOptions());
// cast (lib.rs:378-397): the reference's "deterministic" twin.  On the device both are the same
// computation (the RNG is keyed by pixel/sample, not by visiting order), so this forwards.
Image cast(size_t nx, size_t ny, size_t ns, const camera::Camera& camera, const World& world,
           const CastOptions& opts = CastOptions());
// print_ppm (lib.rs:344-361): byte-identical formatting, to any FILE*.
void print_ppm(const Image& image, std::FILE* out);

// ---------------------------------------------------------------------------------------- scenes
std::vector<object::Box> cornell_box();             // lib.rs:103-166
std::vector<object::Box> cornell_box_with_boxes();  // lib.rs:168-193

struct BuiltScene {
    std::unique_ptr<World> world;
    camera::Camera camera;
    Range exposure{0.f, 1.f};
    Background background = Background::Black;
    std::shared_ptr<const PerlinTables> perlin;
};
// Scenes of src/main.rs:10-319, src/lib.rs:237-319 and benches/scene.rs by name: book1, book1_head,
// cornell, cornell_empty, bench_cornell, final, motion_test, volume_test, simple_light, kitchen_sink.
BuiltScene build_scene(const std::string& name, size_t nx, size_t ny, uint64_t scene_seed, bool use_bvh);

}  // namespace rtiow
