// Host mirror of the crate's Object/Material/Camera surface: bounding boxes, Bvh::new, Camera::look,
// the flattener that turns an object tree into the C ABI's traversal stream, and par_cast/print_ppm
// on top of the C ABI.  Compiled with -ffp-contract=off: scene parameters (box extents, camera
// basis, rotation sines) must come out of the same f32 arithmetic as the reference's.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>

#include "rtiow.hpp"

namespace rtiow {

// ------------------------------------------------------------------------------------------ Vec3
Vec3 operator+(Vec3 a, Vec3 b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
Vec3 operator-(Vec3 a, Vec3 b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
Vec3 operator*(Vec3 a, Vec3 b) { return Vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
Vec3 operator/(Vec3 a, Vec3 b) { return Vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
Vec3 operator*(float s, Vec3 v) { return Vec3::from(s) * v; }                 // vec3.rs:125-132
Vec3 operator+(float s, Vec3 v) { return Vec3(s + v.x, s + v.y, s + v.z); }   // vec3.rs:165-172
Vec3 operator/(Vec3 a, float s) { return Vec3(a.x / s, a.y / s, a.z / s); }   // vec3.rs:145-152
Vec3 operator-(Vec3 a) { return Vec3(-a.x, -a.y, -a.z); }
float Vec3::dot(Vec3 o) const { return (x * o.x + y * o.y) + z * o.z; }       // vec3.rs:43-46
Vec3 Vec3::cross(Vec3 o) const {                                               // vec3.rs:49-55
    return Vec3(y * o.z - z * o.y, -(x * o.z - z * o.x), x * o.y - y * o.x);
}
float Vec3::length() const { return std::sqrt(dot(*this)); }
Vec3 Vec3::into_unit() const { return *this / length(); }

namespace {
inline float fmin2(float a, float b) { return std::fmin(a, b); }  // f32::min ignores NaN
inline float fmax2(float a, float b) { return std::fmax(a, b); }
constexpr float kF32Max = std::numeric_limits<float>::max();
constexpr float kF32Min = std::numeric_limits<float>::lowest();
Vec3 rot(Vec3 p, float s, float c) {  // object.rs:373-379
    return Vec3(p.dot(Vec3(c, 0.f, s)), p.dot(Vec3(0.f, 1.f, 0.f)), p.dot(Vec3(-s, 0.f, c)));
}
}  // namespace

Aabb Aabb::merge(const Aabb& o) const {  // aabb.rs:11-16
    return Aabb{Vec3(fmin2(min.x, o.min.x), fmin2(min.y, o.min.y), fmin2(min.z, o.min.z)),
                Vec3(fmax2(max.x, o.max.x), fmax2(max.y, o.max.y), fmax2(max.z, o.max.z))};
}
std::array<Vec3, 8> Aabb::corners() const {  // aabb.rs:31-43
    std::array<Vec3, 8> c;
    size_t n = 0;
    for (int ix = 0; ix < 2; ++ix)
        for (int iy = 0; iy < 2; ++iy)
            for (int iz = 0; iz < 2; ++iz) c[n++] = Vec3(ix ? max.x : min.x, iy ? max.y : min.y, iz ? max.z : min.z);
    return c;
}

// --------------------------------------------------------------------------------------- SmallRng
SmallRng SmallRng::seed_from_u64(uint64_t state) {
    // rand_core: fill the 16 seed bytes with PCG32 outputs, little-endian
    uint32_t words[4];
    for (uint32_t& w : words) {
        state = state * 6364136223846793005ull + 11634580027462260723ull;
        const uint32_t xorshifted = static_cast<uint32_t>(((state >> 18) ^ state) >> 27);
        const uint32_t r = static_cast<uint32_t>(state >> 59);
        w = (xorshifted >> r) | (xorshifted << ((32u - r) & 31u));
    }
    SmallRng g;
    g.state_ = (static_cast<unsigned __int128>(words[3]) << 96) | (static_cast<unsigned __int128>(words[2]) << 64) |
               (static_cast<unsigned __int128>(words[1]) << 32) | words[0];
    g.state_ |= 1;  // Mcg128Xsl64::new forces the state odd
    return g;
}
uint64_t SmallRng::next_u64() {
    const unsigned __int128 mul = (static_cast<unsigned __int128>(0x2360ED051FC65DA4ull) << 64) | 0x4385DF649FCCF645ull;
    state_ *= mul;
    const unsigned r = static_cast<unsigned>(state_ >> 122);
    const uint64_t x = static_cast<uint64_t>(state_ >> 64) ^ static_cast<uint64_t>(state_);
    return (x >> r) | (x << ((64u - r) & 63u));
}
uint32_t SmallRng::next_u32() { return static_cast<uint32_t>(next_u64()); }
float SmallRng::gen_f32() { return static_cast<float>(next_u32() >> 8) * (1.0f / 16777216.0f); }
Vec3 SmallRng::gen_vec3() {
    const float a = gen_f32();
    const float b = gen_f32();
    const float c = gen_f32();
    return Vec3(a, b, c);
}
float SmallRng::gen_range(float low, float high) {
    if (!(low < high)) throw std::runtime_error("Uniform::sample_single called with low >= high");
    const float scale = high - low, offset = low - scale;
    for (;;) {
        const uint32_t bits = 0x3F800000u | (next_u32() >> 9);
        float v;
        std::memcpy(&v, &bits, 4);
        const float res = v * scale + offset;
        if (res < high) return res;
    }
}
size_t SmallRng::gen_range(size_t low, size_t high) {
    if (!(low < high)) throw std::runtime_error("Uniform::sample_single called with low >= high");
    const uint64_t range = high - low;
    const uint64_t zone = (range << __builtin_clzll(range)) - 1;
    for (;;) {
        const unsigned __int128 wide = static_cast<unsigned __int128>(next_u64()) * range;
        if (static_cast<uint64_t>(wide) <= zone) return low + static_cast<size_t>(wide >> 64);
    }
}

std::shared_ptr<const PerlinTables> PerlinTables::generate(uint64_t scene_seed) {
    SmallRng rng = SmallRng::seed_from_u64(scene_seed ^ 0x5045524C494Eull);
    auto t = std::make_shared<PerlinTables>();
    for (auto& vec : t->vecs) {  // generate_vecs: Vec3::in_unit_sphere  perlin.rs:15-21
        for (;;) {
            const Vec3 v = 2.f * rng.gen_vec3() - Vec3::from(1.f);
            if (v.dot(v) < 1.f) { vec[0] = v.x; vec[1] = v.y; vec[2] = v.z; break; }
        }
    }
    for (auto& p : t->perm) {  // generate_perm  perlin.rs:5-13
        for (int i = 0; i < 256; ++i) p[i] = static_cast<uint8_t>(i);
        for (size_t i = 255; i >= 1; --i) std::swap(p[i], p[rng.gen_range(static_cast<size_t>(0), i)]);
    }
    return t;
}

// --------------------------------------------------------------------------- textures, materials
namespace texture {
Texture constant(Vec3 color) {
    auto n = std::make_shared<Node>();
    n->kind = RTIOW_TEX_CONSTANT;
    n->color = color;
    return n;
}
Texture checker(Texture t0, Texture t1) {
    auto n = std::make_shared<Node>();
    n->kind = RTIOW_TEX_CHECKER;
    n->t0 = std::move(t0);
    n->t1 = std::move(t1);
    return n;
}
Texture perlin(float scale) {
    auto n = std::make_shared<Node>();
    n->kind = RTIOW_TEX_PERLIN;
    n->scale = scale;
    return n;
}
}  // namespace texture

namespace material {
Material Material::Lambertian(texture::Texture a) { Material m; m.kind = RTIOW_MAT_LAMBERTIAN; m.tex = std::move(a); return m; }
Material Material::Metal(Vec3 a, float fuzz) { Material m; m.kind = RTIOW_MAT_METAL; m.albedo = a; m.param = fuzz; return m; }
Material Material::Dielectric(float ri) { Material m; m.kind = RTIOW_MAT_DIELECTRIC; m.param = ri; return m; }
Material Material::DiffuseLight(texture::Texture e, float b) { Material m; m.kind = RTIOW_MAT_DIFFUSE_LIGHT; m.tex = std::move(e); m.param = b; return m; }
Material Material::Isotropic(texture::Texture a) { Material m; m.kind = RTIOW_MAT_ISOTROPIC; m.tex = std::move(a); return m; }
}  // namespace material

// ----------------------------------------------------------------------------------- SceneBuilder
namespace {
template <class T>
std::string bytes_of(const T* p, size_t n) { return std::string(reinterpret_cast<const char*>(p), n * sizeof(T)); }
float as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
}  // namespace

SceneBuilder::SceneBuilder() {
    frames_.push_back(rtiow_frame_t{0, 0});  // frame 0 = world
    frame_index_[std::string()] = 0;
    prefix_stack_.push_back(0);
    desc_.background_kind = RTIOW_BG_BLACK;
}

void SceneBuilder::push_op(uint32_t kind, Vec3 v) {
    rtiow_xform_op_t op{};
    op.kind = kind;
    op.v[0] = v.x; op.v[1] = v.y; op.v[2] = v.z;
    chain_.push_back(op);
}
void SceneBuilder::pop_op() { chain_.pop_back(); }

uint32_t SceneBuilder::intern_frame(size_t n) {
    const std::string key = bytes_of(chain_.data(), n);
    auto& idx = frame_index_;
    auto it = idx.find(key);
    if (it != idx.end()) return it->second;
    const uint32_t id = static_cast<uint32_t>(frames_.size());
    frames_.push_back(rtiow_frame_t{static_cast<uint32_t>(ops_.size()), static_cast<uint32_t>(n)});
    ops_.insert(ops_.end(), chain_.begin(), chain_.begin() + static_cast<std::ptrdiff_t>(n));
    idx[key] = id;
    return id;
}

uint32_t SceneBuilder::intern_texture(const texture::Texture& t) {
    if (!t) throw std::runtime_error("material without a texture");
    rtiow_texture_t rec{};
    rec.kind = t->kind;
    if (t->kind == RTIOW_TEX_CHECKER) {  // children first: child index < parent index
        rec.child0 = intern_texture(t->t0);
        rec.child1 = intern_texture(t->t1);
    } else if (t->kind == RTIOW_TEX_PERLIN) {
        rec.scale = t->scale;
    } else {
        rec.color[0] = t->color.x; rec.color[1] = t->color.y; rec.color[2] = t->color.z;
    }
    const std::string key = bytes_of(&rec, 1);
    auto& idx = texture_index_;
    auto it = idx.find(key);
    if (it != idx.end()) return it->second;
    const uint32_t id = static_cast<uint32_t>(textures_.size());
    textures_.push_back(rec);
    idx[key] = id;
    return id;
}

uint32_t SceneBuilder::intern_material(const material::Material& m) {
    rtiow_material_t rec{};
    rec.kind = m.kind;
    rec.param = m.param;
    if (m.kind == RTIOW_MAT_METAL) {
        rec.albedo[0] = m.albedo.x; rec.albedo[1] = m.albedo.y; rec.albedo[2] = m.albedo.z;
    } else if (m.kind != RTIOW_MAT_DIELECTRIC) {
        rec.tex = intern_texture(m.tex);
    }
    const std::string key = bytes_of(&rec, 1);
    auto& idx = material_index_;
    auto it = idx.find(key);
    if (it != idx.end()) return it->second;
    const uint32_t id = static_cast<uint32_t>(materials_.size());
    materials_.push_back(rec);
    idx[key] = id;
    return id;
}

// Folds the innermost FlipNormals / Translate wrappers (those not already applied by the
// enclosing BBOX frame or medium) into the primitive's own record.  FlipNormals only negates the
// normal and Translate only shifts origin and p, so they commute and the arithmetic is unchanged.
SceneBuilder::Inline SceneBuilder::split_inline(bool allow_offset) const {
    Inline r;
    r.kept = chain_.size();
    const size_t prefix = prefix_stack_.back();
    while (r.kept > prefix) {
        const rtiow_xform_op_t& op = chain_[r.kept - 1];
        if (op.kind == RTIOW_OP_FLIP) {
            r.flip = !r.flip;
        } else if (op.kind == RTIOW_OP_TRANSLATE && allow_offset && !r.has_offset) {
            r.has_offset = true;
            r.offset = Vec3(op.v[0], op.v[1], op.v[2]);
        } else {
            break;
        }
        --r.kept;
    }
    return r;
}

void SceneBuilder::emit_sphere(float radius, const material::Material& m) {
    const Inline in = split_inline(true);
    rtiow_item_t it{};
    it.a[0] = radius;
    it.a_w = RTIOW_ITEM_SPHERE | (intern_frame(in.kept) << 4);
    it.b[0] = in.offset.x; it.b[1] = in.offset.y; it.b[2] = in.offset.z;
    const uint32_t flags = (in.has_offset ? static_cast<uint32_t>(RTIOW_FLAG_HAS_OFFSET) : 0u) | (in.flip ? static_cast<uint32_t>(RTIOW_FLAG_FLIP) : 0u);
    it.b_w = intern_material(m) | (flags << 24);
    items_.push_back(it);
}

void SceneBuilder::emit_rect(int axis, Range r0, Range r1, float k, const material::Material& m) {
    const Inline in = split_inline(false);
    rtiow_item_t it{};
    it.a[0] = k; it.a[1] = r0.start; it.a[2] = r0.end;
    it.a_w = RTIOW_ITEM_RECT | (intern_frame(in.kept) << 4);
    it.b[0] = r1.start; it.b[1] = r1.end;
    const uint32_t flags = (in.flip ? static_cast<uint32_t>(RTIOW_FLAG_FLIP) : 0u) | (static_cast<uint32_t>(axis) << RTIOW_FLAG_AXIS_SHIFT);
    it.b_w = intern_material(m) | (flags << 24);
    items_.push_back(it);
}

size_t SceneBuilder::begin_bbox(const Aabb& box) {
    rtiow_item_t it{};
    it.a[0] = box.min.x; it.a[1] = box.min.y; it.a[2] = box.min.z;
    it.b[0] = box.max.x; it.b[1] = box.max.y; it.b[2] = box.max.z;
    it.a_w = RTIOW_ITEM_BBOX;
    items_.push_back(it);
    return items_.size() - 1;
}
void SceneBuilder::end_bbox(size_t token) {
    items_[token].a_w = RTIOW_ITEM_BBOX | (static_cast<uint32_t>(items_.size()) << 4);  // skip link
}

void SceneBuilder::begin_subtree() {
    if (medium_depth_) {
        // a Bvh as ConstantMedium boundary: its boxes are tested in the medium's frame, so no wrapper may sit between
        // the medium and the Bvh (wrap the medium instead: Translate{ConstantMedium{Bvh}} renders the same)
        if (chain_.size() != prefix_stack_.back())
            throw std::runtime_error("a Bvh used as ConstantMedium boundary must not be wrapped inside the medium; wrap the ConstantMedium instead");
        frame_stack_.push_back(cur_frame_);
        prefix_stack_.push_back(chain_.size());
        return;
    }
    frame_stack_.push_back(cur_frame_);
    if (chain_.size() != prefix_stack_.back()) {  // wrappers since the enclosing frame: boxes live in a new frame
        const uint32_t f = intern_frame(chain_.size());
        rtiow_item_t it{};
        it.a_w = RTIOW_ITEM_SET_FRAME | (f << 4);
        items_.push_back(it);
        cur_frame_ = f;
    }
    prefix_stack_.push_back(chain_.size());
}
void SceneBuilder::end_subtree() {
    prefix_stack_.pop_back();
    const uint32_t prev = frame_stack_.back();
    frame_stack_.pop_back();
    if (prev != cur_frame_) {
        rtiow_item_t it{};
        it.a_w = RTIOW_ITEM_SET_FRAME | (prev << 4);
        items_.push_back(it);
        cur_frame_ = prev;
    }
}

void SceneBuilder::begin_medium(float density, const material::Material& m, uint32_t medium_id) {
    if (medium_depth_) throw std::runtime_error("a ConstantMedium inside a ConstantMedium boundary is not supported");
    rtiow_item_t it{};
    it.a[0] = density;
    it.a[1] = as_float(medium_id);
    it.a_w = RTIOW_ITEM_MEDIUM | (intern_frame(chain_.size()) << 4);
    it.b_w = intern_material(m);
    items_.push_back(it);
    medium_item_ = items_.size() - 1;
    prefix_stack_.push_back(chain_.size());
    ++medium_depth_;
}
void SceneBuilder::end_medium() {
    --medium_depth_;
    prefix_stack_.pop_back();
    if (items_.size() == medium_item_ + 1) throw std::runtime_error("ConstantMedium boundary flattened to nothing");
    items_[medium_item_].a[2] = as_float(static_cast<uint32_t>(items_.size()));  // index after the boundary run
}

void SceneBuilder::set_background(uint32_t kind, Vec3 c0, Vec3 c1) {
    desc_.background_kind = kind;
    desc_.background_c0[0] = c0.x; desc_.background_c0[1] = c0.y; desc_.background_c0[2] = c0.z;
    desc_.background_c1[0] = c1.x; desc_.background_c1[1] = c1.y; desc_.background_c1[2] = c1.z;
}

const rtiow_scene_desc_t& SceneBuilder::finish() {
    if (!finished_) {
        items_.push_back(rtiow_item_t{});  // RTIOW_ITEM_END
        finished_ = true;
    }
    desc_.abi_version = RTIOW_B200_ABI_VERSION;
    desc_.n_items = static_cast<uint32_t>(items_.size());
    desc_.items = items_.data();
    desc_.n_frames = static_cast<uint32_t>(frames_.size());
    desc_.frames = frames_.data();
    desc_.n_ops = static_cast<uint32_t>(ops_.size());
    desc_.ops = ops_.data();
    desc_.n_materials = static_cast<uint32_t>(materials_.size());
    desc_.materials = materials_.data();
    desc_.n_textures = static_cast<uint32_t>(textures_.size());
    desc_.textures = textures_.data();
    desc_.perlin_vecs = perlin_ ? &perlin_->vecs[0][0] : nullptr;
    desc_.perlin_perm = perlin_ ? &perlin_->perm[0][0] : nullptr;
    return desc_;
}

// ---------------------------------------------------------------------------------------- objects
namespace object {

Aabb Sphere::bounding_box(Range) const { return Aabb{-Vec3::from(radius), Vec3::from(radius)}; }  // object.rs:113-118
void Sphere::flatten(SceneBuilder& b) const { b.emit_sphere(radius, material); }

Aabb Rect::bounding_box(Range) const {  // object.rs:220-233
    const int a = static_cast<int>(orthogonal_to);
    const int o1 = a == 0 ? 1 : 0, o2 = a == 2 ? 1 : 2;  // "other two" alphabetical (object.rs:153-181)
    Vec3 mn, mx;
    mn[a] = k - 0.0001f;
    mx[a] = k + 0.0001f;
    mn[o1] = range0.start; mx[o1] = range0.end;
    mn[o2] = range1.start; mx[o2] = range1.end;
    return Aabb{mn, mx};
}
void Rect::flatten(SceneBuilder& b) const { b.emit_rect(static_cast<int>(orthogonal_to), range0, range1, k, material); }

Aabb FlipNormals::bounding_box(Range e) const { return inner->bounding_box(e); }
void FlipNormals::flatten(SceneBuilder& b) const {
    b.push_op(RTIOW_OP_FLIP, Vec3());
    inner->flatten(b);
    b.pop_op();
}

Aabb Translate::bounding_box(Range e) const {  // object.rs:285-291
    const Aabb bb = object->bounding_box(e);
    return Aabb{bb.min + offset, bb.max + offset};
}
void Translate::flatten(SceneBuilder& b) const {
    b.push_op(RTIOW_OP_TRANSLATE, offset);
    object->flatten(b);
    b.pop_op();
}

Aabb Scale::bounding_box(Range e) const {  // object.rs:321-327
    const Aabb bb = object->bounding_box(e);
    return Aabb{bb.min * factor, bb.max * factor};
}
void Scale::flatten(SceneBuilder& b) const {
    b.push_op(RTIOW_OP_SCALE, factor);
    object->flatten(b);
    b.pop_op();
}

Aabb RotateY::bounding_box(Range e) const {  // object.rs:372-389
    Vec3 mn = Vec3::from(kF32Max), mx = Vec3::from(kF32Min);
    for (const Vec3& c : object->bounding_box(e).corners()) {
        const Vec3 r = rot(c, sin_theta, cos_theta);
        mn = Vec3(fmin2(mn.x, r.x), fmin2(mn.y, r.y), fmin2(mn.z, r.z));
        mx = Vec3(fmax2(mx.x, r.x), fmax2(mx.y, r.y), fmax2(mx.z, r.z));
    }
    return Aabb{mn, mx};
}
void RotateY::flatten(SceneBuilder& b) const {
    b.push_op(RTIOW_OP_ROTATE_Y, Vec3(sin_theta, cos_theta, 0.f));
    object->flatten(b);
    b.pop_op();
}
Box rotate_y(float degrees, Box object) {  // object.rs:477-484
    const float radians = degrees * 3.14159265358979323846f / 180.f;
    return std::make_unique<RotateY>(std::move(object), std::sin(radians), std::cos(radians));
}

Aabb And::bounding_box(Range e) const { return first->bounding_box(e).merge(second->bounding_box(e)); }
void And::flatten(SceneBuilder& b) const {  // hit order: .0 then .1 with the tightened range (object.rs:403-409)
    first->flatten(b);
    second->flatten(b);
}

Box rect_prism(Vec3 p0, Vec3 p1, const material::Material& m) {  // object.rs:420-473
    auto R = [&](StaticAxis a, float a0, float a1, float b0, float b1, float k) -> Box {
        return std::make_unique<Rect>(a, Range{a0, a1}, Range{b0, b1}, k, m);
    };
    auto F = [](Box o) -> Box { return std::make_unique<FlipNormals>(std::move(o)); };
    auto A = [](Box l, Box r) -> Box { return std::make_unique<And>(std::move(l), std::move(r)); };
    return A(A(R(StaticZ, p0.x, p1.x, p0.y, p1.y, p1.z),
               A(R(StaticY, p0.x, p1.x, p0.z, p1.z, p1.y), R(StaticX, p0.y, p1.y, p0.z, p1.z, p1.x))),
             A(F(R(StaticZ, p0.x, p1.x, p0.y, p1.y, p0.z)),
               A(F(R(StaticY, p0.x, p1.x, p0.z, p1.z, p0.y)), F(R(StaticX, p0.y, p1.y, p0.z, p1.z, p0.x)))));
}

Aabb LinearMove::bounding_box(Range e) const {  // object.rs:514-527
    const Aabb bb = object->bounding_box(e);
    const Aabb s{bb.min + e.start * motion, bb.max + e.start * motion};
    const Aabb f{bb.min + e.end * motion, bb.max + e.end * motion};
    return s.merge(f);
}
void LinearMove::flatten(SceneBuilder& b) const {
    b.push_op(RTIOW_OP_LINEAR_MOVE, motion);
    object->flatten(b);
    b.pop_op();
}

Aabb ConstantMedium::bounding_box(Range e) const { return boundary->bounding_box(e); }
void ConstantMedium::flatten(SceneBuilder& b) const {
    b.begin_medium(density, material, medium_id);
    boundary->flatten(b);
    b.end_medium();
}

}  // namespace object

// -------------------------------------------------------------------------------------------- Bvh
namespace bvh {

Bvh::Bvh(std::vector<object::Box> objs, Range exposure) {
    if (objs.empty()) throw std::runtime_error("Can't create a BVH from zero objects.");
    // the axis with the greatest extent over these objects (bvh.rs:27-46)
    float extent[3];
    for (int axis = 0; axis < 3; ++axis) {
        float lo = kF32Max, hi = kF32Min;
        for (const auto& o : objs) {
            const Aabb bb = o->bounding_box(exposure);
            lo = fmin2(lo, fmin2(bb.min[axis], bb.max[axis]));
            hi = fmax2(hi, fmax2(bb.min[axis], bb.max[axis]));
        }
        extent[axis] = hi - lo;
        if (std::isnan(extent[axis])) throw std::runtime_error("called `Option::unwrap()` on a `None` value (NaN extent in Bvh::new)");
    }
    int axis = 0;
    for (int a = 1; a < 3; ++a)
        if (extent[a] > extent[axis]) axis = a;
    // sort by centroid*2 on that axis (bvh.rs:51-57)
    std::vector<std::pair<float, size_t>> order;
    order.reserve(objs.size());
    for (size_t i = 0; i < objs.size(); ++i) {
        const Aabb bb = objs[i]->bounding_box(exposure);
        const float key = bb.min[axis] + bb.max[axis];
        if (std::isnan(key)) throw std::runtime_error("called `Option::unwrap()` on a `None` value (NaN centroid in Bvh::new)");
        order.emplace_back(key, i);
    }
    std::stable_sort(order.begin(), order.end(), [](const auto& l, const auto& r) { return l.first < r.first; });

    if (objs.size() == 1) {  // bvh.rs:61-65
        bounding_box_ = objs[0]->bounding_box(exposure);
        size_ = 1;
        leaf_ = std::move(objs[0]);
        return;
    }
    const size_t half = objs.size() / 2;  // bvh.rs:68-72: right = drain(len/2..), built first
    std::vector<object::Box> lo_half, hi_half;
    for (size_t r = 0; r < order.size(); ++r) (r < half ? lo_half : hi_half).push_back(std::move(objs[order[r].second]));
    right_ = std::make_unique<Bvh>(std::move(hi_half), exposure);
    left_ = std::make_unique<Bvh>(std::move(lo_half), exposure);
    bounding_box_ = left_->bounding_box_.merge(right_->bounding_box_);
    size_ = left_->size_ + right_->size_;
}

size_t Bvh::node_count() const { return leaf_ ? 1 : 1 + left_->node_count() + right_->node_count(); }

void Bvh::flatten_node(SceneBuilder& b) const {  // Bvh::hit visiting order (bvh.rs:85-120)
    const size_t tok = b.begin_bbox(bounding_box_);
    if (leaf_) {
        leaf_->flatten(b);
    } else {
        left_->flatten_node(b);
        right_->flatten_node(b);
    }
    b.end_bbox(tok);
}
void Bvh::flatten(SceneBuilder& b) const {
    b.begin_subtree();
    flatten_node(b);
    b.end_subtree();
}
std::unique_ptr<Bvh> from_scene(std::vector<object::Box> scene, Range exposure) {
    return std::make_unique<Bvh>(std::move(scene), exposure);
}

}  // namespace bvh

// ----------------------------------------------------------------------------------------- Camera
namespace camera {
Camera Camera::look(Vec3 look_from, Vec3 look_at, Vec3 up, float fov, float aspect, float aperture, float focus_dist,
                    Range exposure) {
    Camera c;
    c.lens_radius = aperture / 2.f;
    const float theta = fov * 3.14159265358979323846f / 180.f;
    const float half_height = std::tan(theta / 2.f);
    const float half_width = aspect * half_height;
    c.origin = look_from;
    const Vec3 w = (look_from - look_at).into_unit();
    c.u = up.cross(w).into_unit();
    c.v = w.cross(c.u);
    c.lower_left_corner = c.origin - half_width * focus_dist * c.u - half_height * focus_dist * c.v - focus_dist * w;
    c.horizontal = 2.f * half_width * focus_dist * c.u;
    c.vertical = 2.f * half_height * focus_dist * c.v;
    c.exposure = exposure;
    return c;
}
rtiow_camera_t Camera::to_repr_c() const {
    rtiow_camera_t r{};
    auto put = [](float* d, Vec3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; };
    put(r.origin, origin); put(r.lower_left_corner, lower_left_corner); put(r.horizontal, horizontal);
    put(r.vertical, vertical); put(r.u, u); put(r.v, v);
    r.lens_radius = lens_radius;
    r.time0 = exposure.start;
    r.time1 = exposure.end;
    return r;
}
}  // namespace camera

// ------------------------------------------------------------------------------------------ World
void World::flatten(SceneBuilder& b) const {
    if (bvh_) {
        bvh_->flatten(b);  // lib.rs:51-55
    } else {
        for (const auto& o : list_) o->flatten(b);  // lib.rs:40-45: in order, shrinking `nearest`
    }
}

Image par_cast(size_t nx, size_t ny, size_t ns, const camera::Camera& cam, const World& world, const CastOptions& opts) {
    SceneBuilder b;
    b.set_perlin(opts.perlin);
    b.set_background(static_cast<uint32_t>(opts.background), Vec3(1.f, 1.f, 1.f), Vec3(0.5f, 0.7f, 1.0f));
    world.flatten(b);
    rtiow_scene_t* scene = nullptr;
    if (rtiow_b200_scene_create(&b.finish(), opts.device, &scene) != RTIOW_OK)
        throw std::runtime_error(std::string("rtiow_b200_scene_create: ") + rtiow_b200_last_error());
    Image img;
    img.nx = nx;
    img.ny = ny;
    img.rgb.resize(nx * ny * 3);
    const rtiow_camera_t c = cam.to_repr_c();
    const int rc = rtiow_b200_render(scene, &c, static_cast<uint32_t>(nx), static_cast<uint32_t>(ny), static_cast<uint32_t>(ns),
                                     opts.seed, img.rgb.data());
    const std::string err = rc ? rtiow_b200_last_error() : "";
    rtiow_b200_scene_destroy(scene);
    if (rc) throw std::runtime_error("rtiow_b200_render: " + err);
    return img;
}

Image cast(size_t nx, size_t ny, size_t ns, const camera::Camera& cam, const World& world, const CastOptions& opts) {
    return par_cast(nx, ny, ns, cam, world, opts);
}

void print_ppm(const Image& image, std::FILE* out) {  // lib.rs:344-361
    std::fprintf(out, "P3\n%zu %zu\n255\n", image.nx, image.ny);
    auto to_u8 = [](float x) {
        const float v = 255.99f * x;
        int i;
        if (std::isnan(v)) i = 0;                          // Rust `as i32`: NaN -> 0, saturating
        else if (v >= 2147483648.f) i = 2147483647;
        else if (v <= -2147483648.f) i = -2147483647 - 1;
        else i = static_cast<int>(v);
        return std::min(std::max(i, 0), 255);
    };
    for (size_t p = 0; p < image.nx * image.ny; ++p) {
        const float* c = &image.rgb[3 * p];
        std::fprintf(out, "%d %d %d\n", to_u8(std::sqrt(c[0])), to_u8(std::sqrt(c[1])), to_u8(std::sqrt(c[2])));
    }
}

}  // namespace rtiow
