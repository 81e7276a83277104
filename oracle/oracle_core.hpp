// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.hpp for the rules).  PARITY UNPINNED.
//
// Tree-walking CPU restatement of the reference's render path: Ray, Aabb, the Object trait and
// every implementor, Bvh, Material, Texture closures, Perlin noise, Camera, hit_top, color, cast.
// It walks the object TREE (Box<dyn Object> style), never the product's flattened buffers, so it
// is independent of the product's flattener and kernel.
#pragma once
#include <algorithm>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "oracle_math.hpp"

namespace oracle {

// ---------------------------------------------------------------------------------------------
// Work counters (algorithmic bytes are defined on the reference's own traversal).
// ---------------------------------------------------------------------------------------------
struct Counters {
    uint64_t samples = 0, segments = 0, node_tests = 0, sphere_tests = 0, rect_tests = 0,
             medium_evals = 0, draws = 0, max_segments = 0;
    void add(const Counters& o) {
        samples += o.samples; segments += o.segments; node_tests += o.node_tests;
        sphere_tests += o.sphere_tests; rect_tests += o.rect_tests; medium_evals += o.medium_evals;
        draws += o.draws; max_segments = std::max(max_segments, o.max_segments);
    }
};

// ---------------------------------------------------------------------------------------------
// Per-sample RNG context.  Where the reference pulls the next f32 from a sequential generator,
// the contract pulls a word addressed by (seed; pixel, sample, bounce, purpose, index):
//   key     = (seed_lo, seed_hi)
//   counter = (pixel = y*nx + x with y counted from the bottom as in lib.rs:368-369,
//              sample, (bounce << 16) | purpose, index)
//   purpose 0 CAMERA  index 0: w0 -> u jitter (lib.rs:368), w1 -> v jitter (lib.rs:369),
//                              w2,w3 -> shutter attempts 0,1 (camera.rs:55); attempt k>=2 is word
//                              (k-2)%4 of index 1+(k-2)/4
//   purpose 1 LENS    attempt k of in_unit_disc (camera.rs:53, vec3.rs:32-39) = words
//                              2(k%2), 2(k%2)+1 of index k/2
//   purpose 2 SCATTER attempt k of in_unit_sphere (vec3.rs:19-26) = words 0,1,2 of index k;
//                              Dielectric's single draw (material.rs:97) = word 0 of index 0
//   purpose 16+m MEDIUM  ConstantMedium #m's draw (object.rs:562) = word 0 of index 0
// so the number a consumer gets never depends on traversal order.
// ---------------------------------------------------------------------------------------------
enum : uint32_t { PURPOSE_CAMERA = 0, PURPOSE_LENS = 1, PURPOSE_SCATTER = 2, PURPOSE_MEDIUM0 = 16 };

struct PathRng {
    uint32_t key[2];
    uint32_t pixel = 0, sample = 0, bounce = 0;
    Counters* counters = nullptr;

    PathRng(uint64_t seed, uint32_t pixel_, uint32_t sample_, Counters* c)
        : pixel(pixel_), sample(sample_), counters(c) {
        key[0] = static_cast<uint32_t>(seed);
        key[1] = static_cast<uint32_t>(seed >> 32);
    }
    void block(uint32_t purpose, uint32_t index, uint32_t out[4]) const {
        uint32_t ctr[4] = {pixel, sample, (bounce << 16) | purpose, index};
        philox4x32_10(key, ctr, out);
    }
    void count(unsigned n) const { if (counters) counters->draws += n; }
};

// vec3.rs:19-26.  `2. * rng.gen::<Vec3>() - Vec3::from(1.)`, accept iff v.dot(v) < 1.
inline Vec3 in_unit_sphere(const PathRng& rng) {
    for (uint32_t k = 0;; ++k) {
        uint32_t w[4];
        rng.block(PURPOSE_SCATTER, k, w);
        rng.count(3);
        Vec3 g(u32_to_unit_f32(w[0]), u32_to_unit_f32(w[1]), u32_to_unit_f32(w[2]));
        Vec3 v = 2.f * g - splat(1.f);
        if (dot(v, v) < 1.f) return v;
    }
}

// vec3.rs:32-39.  `2. * Vec3(gen, gen, 0.) - Vec3(1., 1., 0.)`
inline Vec3 in_unit_disc(const PathRng& rng) {
    for (uint32_t k = 0;; ++k) {
        uint32_t w[4];
        rng.block(PURPOSE_LENS, k / 2, w);
        rng.count(2);
        unsigned o = 2 * (k % 2);
        Vec3 v = 2.f * Vec3(u32_to_unit_f32(w[o]), u32_to_unit_f32(w[o + 1]), 0.f) - Vec3(1.f, 1.f, 0.f);
        if (dot(v, v) < 1.f) return v;
    }
}

// ---------------------------------------------------------------------------------------------
// Ray — src/ray.rs
// ---------------------------------------------------------------------------------------------
struct Ray {
    Vec3 origin, direction;
    float time = 0.f;
    Vec3 point_at_parameter(float t) const { return origin + t * direction; }  // ray.rs:15-17
};

// ---------------------------------------------------------------------------------------------
// Aabb — src/aabb.rs
// ---------------------------------------------------------------------------------------------
struct Aabb {
    Vec3 min, max;
    Aabb merge(const Aabb& o) const {  // aabb.rs:11-16
        return Aabb{Vec3(rmin(min.x, o.min.x), rmin(min.y, o.min.y), rmin(min.z, o.min.z)),
                    Vec3(rmax(max.x, o.max.x), rmax(max.y, o.max.y), rmax(max.z, o.max.z))};
    }
    bool hit(const Ray& ray, float t_start, float t_end) const {  // aabb.rs:18-29
        Vec3 inv_d(1.f / ray.direction.x, 1.f / ray.direction.y, 1.f / ray.direction.z);
        Vec3 t0 = (min - ray.origin) * inv_d;
        Vec3 t1 = (max - ray.origin) * inv_d;
        Vec3 n0(inv_d.x < 0.f ? t1.x : t0.x, inv_d.y < 0.f ? t1.y : t0.y, inv_d.z < 0.f ? t1.z : t0.z);
        Vec3 n1(inv_d.x < 0.f ? t0.x : t1.x, inv_d.y < 0.f ? t0.y : t1.y, inv_d.z < 0.f ? t0.z : t1.z);
        float start = rmax(t_start, rmax(rmax(n0.x, n0.y), n0.z));
        float end = rmin(t_end, rmin(rmin(n1.x, n1.y), n1.z));
        return end > start;
    }
    void corners(Vec3 out[8]) const {  // aabb.rs:31-43 (x outermost, z innermost)
        int n = 0;
        for (int ix = 0; ix < 2; ++ix)
            for (int iy = 0; iy < 2; ++iy)
                for (int iz = 0; iz < 2; ++iz)
                    out[n++] = Vec3(ix == 0 ? min.x : max.x, iy == 0 ? min.y : max.y, iz == 0 ? min.z : max.z);
    }
};

// ---------------------------------------------------------------------------------------------
// Perlin — src/perlin.rs.  The reference fills the tables from thread_rng() in a lazy_static
// (perlin.rs:24-29), i.e. differently on every run; here they are scene data.
// ---------------------------------------------------------------------------------------------
struct PerlinTables {
    Vec3 vecs[256];
    uint8_t perm_x[256], perm_y[256], perm_z[256];
};

inline float trilinear_interp(const Vec3 corners[2][2][2], Vec3 uvw) {  // perlin.rs:31-47
    float accum = 0.f;
    Vec3 uvw3 = uvw * uvw * (splat(3.f) - 2.f * uvw);
    Vec3 uvw3_inv = splat(1.f) - uvw3;
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            for (int k = 0; k < 2; ++k) {
                Vec3 ijk(static_cast<float>(i), static_cast<float>(j), static_cast<float>(k));
                float weight = dot(corners[i][j][k], uvw - ijk);
                Vec3 ijk_inv = splat(1.f) - ijk;
                Vec3 m = ijk * uvw3 + ijk_inv * uvw3_inv;
                accum = accum + ((m.x * m.y) * m.z) * weight;
            }
    return accum;
}

inline float perlin_noise(const PerlinTables& tb, Vec3 p) {  // perlin.rs:49-64
    Vec3 ijk(std::floor(p.x), std::floor(p.y), std::floor(p.z));
    Vec3 uvw = p - ijk;
    Vec3 corners[2][2][2];
    for (int di = 0; di < 2; ++di)
        for (int dj = 0; dj < 2; ++dj)
            for (int dk = 0; dk < 2; ++dk) {
                // wrapping add is irrelevant: `& 255` only looks at the low byte
                uint8_t ix = tb.perm_x[static_cast<uint32_t>(f32_as_i32(ijk.x) + di) & 255u];
                uint8_t iy = tb.perm_y[static_cast<uint32_t>(f32_as_i32(ijk.y) + dj) & 255u];
                uint8_t iz = tb.perm_z[static_cast<uint32_t>(f32_as_i32(ijk.z) + dk) & 255u];
                corners[di][dj][dk] = tb.vecs[ix ^ iy ^ iz];
            }
    return trilinear_interp(corners, uvw);
}

inline float perlin_turb(const PerlinTables& tb, Vec3 p, int depth) {  // perlin.rs:66-75
    float accum = 0.f, weight = 1.f;
    for (int i = 0; i < depth; ++i) {
        accum += weight * perlin_noise(tb, p);
        weight *= 0.5f;
        p = 2.f * p;
    }
    return std::fabs(accum);
}

// ---------------------------------------------------------------------------------------------
// Texture — src/texture.rs: Arc<dyn Fn(Vec3) -> Vec3>
// ---------------------------------------------------------------------------------------------
using Texture = std::shared_ptr<std::function<Vec3(Vec3)>>;

inline Texture tex_constant(Vec3 color) {  // texture.rs:8-10
    return std::make_shared<std::function<Vec3(Vec3)>>([color](Vec3) { return color; });
}
inline Texture tex_checker(Texture t0, Texture t1) {  // texture.rs:12-21
    return std::make_shared<std::function<Vec3(Vec3)>>([t0, t1](Vec3 p) {
        Vec3 q = 10.f * p;
        float s = (sin_f32(q.x) * sin_f32(q.y)) * sin_f32(q.z);
        return s < 0.f ? (*t1)(p) : (*t0)(p);
    });
}
inline Texture tex_perlin(std::shared_ptr<const PerlinTables> tb, float scale) {  // texture.rs:23-26
    return std::make_shared<std::function<Vec3(Vec3)>>(
        [tb, scale](Vec3 p) { return splat(perlin_turb(*tb, scale * p, 7)); });
}

// ---------------------------------------------------------------------------------------------
// Material — src/material.rs
// ---------------------------------------------------------------------------------------------
struct HitRecord;

struct Material {
    enum Kind { Lambertian, Metal, Dielectric, DiffuseLight, Isotropic } kind = Lambertian;
    Texture albedo;          // Lambertian / Isotropic: albedo; DiffuseLight: emission
    Vec3 metal_albedo;       // Metal
    float fuzz = 0.f;        // Metal
    float ref_idx = 1.f;     // Dielectric
    float brightness = 0.f;  // DiffuseLight

    static Material lambertian(Texture a) { Material m; m.kind = Lambertian; m.albedo = std::move(a); return m; }
    static Material metal(Vec3 a, float fuzz) { Material m; m.kind = Metal; m.metal_albedo = a; m.fuzz = fuzz; return m; }
    static Material dielectric(float ri) { Material m; m.kind = Dielectric; m.ref_idx = ri; return m; }
    static Material diffuse_light(Texture e, float b) { Material m; m.kind = DiffuseLight; m.albedo = std::move(e); m.brightness = b; return m; }
    static Material isotropic(Texture a) { Material m; m.kind = Isotropic; m.albedo = std::move(a); return m; }

    bool scatter(const Ray& ray, const HitRecord& hit, const PathRng& rng, Ray& scattered, Vec3& attenuation) const;
    Vec3 emitted(Vec3 p) const {  // material.rs:120-128
        if (kind == DiffuseLight) return brightness * (*albedo)(p);
        return Vec3();
    }
};

struct HitRecord {  // object.rs:61-71
    float t = 0.f;
    Vec3 p, normal;
    const Material* material = nullptr;
};

inline float schlick(float cos, float ref_idx) {  // material.rs:142-146
    float r0 = (1.f - ref_idx) / (1.f + ref_idx);
    r0 = r0 * r0;
    return r0 + (1.f - r0) * pow5_f32(1.f - cos);
}

inline bool Material::scatter(const Ray& ray, const HitRecord& hit, const PathRng& rng, Ray& scattered,
                              Vec3& attenuation) const {  // material.rs:55-118
    switch (kind) {
        case Lambertian: {
            Vec3 target = hit.p + hit.normal + in_unit_sphere(rng);
            scattered.origin = hit.p;
            scattered.direction = target - hit.p;
            scattered.time = ray.time;
            attenuation = (*albedo)(hit.p);
            return true;
        }
        case Metal: {
            Vec3 refl = reflect(into_unit(ray.direction), hit.normal);
            scattered.origin = hit.p;
            scattered.direction = refl + fuzz * in_unit_sphere(rng);
            scattered.time = ray.time;
            if (dot(scattered.direction, hit.normal) > 0.f) {
                attenuation = metal_albedo;
                return true;
            }
            return false;
        }
        case Dielectric: {
            Vec3 outward_normal;
            float ni_over_nt, cosine;
            if (dot(ray.direction, hit.normal) > 0.f) {
                outward_normal = -hit.normal;
                ni_over_nt = ref_idx;
                cosine = ref_idx * dot(ray.direction, hit.normal) / length(ray.direction);
            } else {
                outward_normal = hit.normal;
                ni_over_nt = 1.0f / ref_idx;
                cosine = -dot(ray.direction, hit.normal) / length(ray.direction);
            }
            Vec3 direction;
            bool refracted = refract(ray.direction, outward_normal, ni_over_nt, direction);
            if (refracted) {  // .filter(|_| rng.gen::<f32>() >= schlick(..)): drawn only if Some
                uint32_t w[4];
                rng.block(PURPOSE_SCATTER, 0, w);
                rng.count(1);
                if (!(u32_to_unit_f32(w[0]) >= schlick(cosine, ref_idx))) refracted = false;
            }
            if (!refracted) direction = reflect(ray.direction, hit.normal);
            attenuation = splat(1.f);
            scattered.origin = hit.p;
            scattered.direction = direction;
            scattered.time = ray.time;
            return true;
        }
        case DiffuseLight:
            return false;
        case Isotropic: {
            scattered.origin = hit.p;
            scattered.direction = in_unit_sphere(rng);
            scattered.time = ray.time;
            attenuation = (*albedo)(hit.p);
            return true;
        }
    }
    return false;
}

// ---------------------------------------------------------------------------------------------
// Object trait and implementors — src/object.rs
// ---------------------------------------------------------------------------------------------
constexpr float F32_MAX = std::numeric_limits<float>::max();      // std::f32::MAX
constexpr float F32_MIN = std::numeric_limits<float>::lowest();   // std::f32::MIN (most negative)

struct Object {  // object.rs:15-40
    virtual ~Object() = default;
    virtual bool hit(const Ray& ray, float t_start, float t_end, const PathRng& rng, HitRecord& rec) const = 0;
    virtual Aabb bounding_box(float e0, float e1) const = 0;
};
using ObjectBox = std::unique_ptr<Object>;  // Box<dyn Object>

struct Sphere : Object {  // object.rs:74-119
    float radius;
    Material material;
    Sphere(float r, Material m) : radius(r), material(std::move(m)) {}
    bool hit(const Ray& ray, float t_start, float t_end, const PathRng& rng, HitRecord& rec) const override {
        if (rng.counters) rng.counters->sphere_tests++;
        float a = dot(ray.direction, ray.direction);
        float b = dot(ray.origin, ray.direction);
        float c = dot(ray.origin, ray.origin) - radius * radius;
        float discriminant = b * b - a * c;
        if (discriminant > 0.f) {
            const float ts[2] = {(-b - std::sqrt(discriminant)) / a, (-b + std::sqrt(discriminant)) / a};
            for (float t : ts) {
                if (t < t_end && t >= t_start) {
                    Vec3 p = ray.point_at_parameter(t);
                    rec.t = t;
                    rec.p = p;
                    rec.normal = p / radius;
                    rec.material = &material;
                    return true;
                }
            }
        }
        return false;
    }
    Aabb bounding_box(float, float) const override { return Aabb{-splat(radius), splat(radius)}; }
};

// Rect<A> — object.rs:131-234.  axis: 0=X,1=Y,2=Z; the "other two" are alphabetical (:153-181).
struct Rect : Object {
    int axis;
    float r0s, r0e, r1s, r1e, k;
    Material material;
    Rect(int axis_, float r0s_, float r0e_, float r1s_, float r1e_, float k_, Material m)
        : axis(axis_), r0s(r0s_), r0e(r0e_), r1s(r1s_), r1e(r1e_), k(k_), material(std::move(m)) {}
    int other1() const { return axis == 0 ? 1 : 0; }
    int other2() const { return axis == 2 ? 1 : 2; }
    bool hit(const Ray& ray, float t_start, float t_end, const PathRng& rng, HitRecord& rec) const override {
        if (rng.counters) rng.counters->rect_tests++;
        float t = (k - ray.origin[axis]) / ray.direction[axis];
        if (t < t_start || t >= t_end) return false;
        float x = ray.origin[other1()] + t * ray.direction[other1()];
        float y = ray.origin[other2()] + t * ray.direction[other2()];
        if (x < r0s || x >= r0e || y < r1s || y >= r1e) return false;
        Vec3 normal;
        normal[axis] = 1.f;
        rec.t = t;
        rec.p = ray.point_at_parameter(t);
        rec.material = &material;
        rec.normal = normal;
        return true;
    }
    Aabb bounding_box(float, float) const override {  // object.rs:220-233
        Vec3 mn, mx;
        mn[axis] = k - 0.0001f;
        mx[axis] = k + 0.0001f;
        mn[other1()] = r0s;
        mx[other1()] = r0e;
        mn[other2()] = r1s;
        mx[other2()] = r1e;
        return Aabb{mn, mx};
    }
};

struct FlipNormals : Object {  // object.rs:238-258
    ObjectBox inner;
    explicit FlipNormals(ObjectBox o) : inner(std::move(o)) {}
    bool hit(const Ray& ray, float t_start, float t_end, const PathRng& rng, HitRecord& rec) const override {
        if (!inner->hit(ray, t_start, t_end, rng, rec)) return false;
        rec.normal = -rec.normal;
        return true;
    }
    Aabb bounding_box(float e0, float e1) const override { return inner->bounding_box(e0, e1); }
};

struct Translate : Object {  // object.rs:261-292
    Vec3 offset;
    ObjectBox object;
    Translate(Vec3 off, ObjectBox o) : offset(off), object(std::move(o)) {}
    bool hit(const Ray& ray, float t_start, float t_end, const PathRng& rng, HitRecord& rec) const override {
        Ray t_ray = ray;
        t_ray.origin = ray.origin - offset;
        if (!object->hit(t_ray, t_start, t_end, rng, rec)) return false;
        rec.p = rec.p + offset;
        return true;
    }
    Aabb bounding_box(float e0, float e1) const override {
        Aabb b = object->bounding_box(e0, e1);
        return Aabb{b.min + offset, b.max + offset};
    }
};

struct Scale : Object {  // object.rs:295-328
    Vec3 factor;
    ObjectBox object;
    Scale(Vec3 f, ObjectBox o) : factor(f), object(std::move(o)) {}
    bool hit(const Ray& ray, float t_start, float t_end, const PathRng& rng, HitRecord& rec) const override {
        Ray t_ray = ray;
        t_ray.origin = ray.origin / factor;
        t_ray.direction = ray.direction / factor;
        if (!object->hit(t_ray, t_start, t_end, rng, rec)) return false;
        rec.p = rec.p * factor;
        rec.normal = rec.normal / factor;
        return true;
    }
    Aabb bounding_box(float e0, float e1) const override {
        Aabb b = object->bounding_box(e0, e1);
        return Aabb{b.min * factor, b.max * factor};
    }
};

inline Vec3 rot_y(Vec3 p, float sin_theta, float cos_theta) {  // object.rs:349-355 / :373-379
    return Vec3(dot(p, Vec3(cos_theta, 0.f, sin_theta)), dot(p, Vec3(0.f, 1.f, 0.f)),
                dot(p, Vec3(-sin_theta, 0.f, cos_theta)));
}

struct RotateY : Object {  // object.rs:335-390
    ObjectBox object;
    float sin_theta, cos_theta;
    RotateY(ObjectBox o, float s, float c) : object(std::move(o)), sin_theta(s), cos_theta(c) {}
    bool hit(const Ray& ray, float t_start, float t_end, const PathRng& rng, HitRecord& rec) const override {
        Ray rot_ray = ray;
        rot_ray.origin = rot_y(ray.origin, -sin_theta, cos_theta);
        rot_ray.direction = rot_y(ray.direction, -sin_theta, cos_theta);
        if (!object->hit(rot_ray, t_start, t_end, rng, rec)) return false;
        rec.p = rot_y(rec.p, sin_theta, cos_theta);
        rec.normal = rot_y(rec.normal, sin_theta, cos_theta);
        return true;
    }
    Aabb bounding_box(float e0, float e1) const override {
        Vec3 c[8];
        object->bounding_box(e0, e1).corners(c);
        Vec3 mn = splat(F32_MAX), mx = splat(F32_MIN);
        for (const Vec3& corner : c) {
            Vec3 r = rot_y(corner, sin_theta, cos_theta);
            mn = Vec3(rmin(mn.x, r.x), rmin(mn.y, r.y), rmin(mn.z, r.z));
            mx = Vec3(rmax(mx.x, r.x), rmax(mx.y, r.y), rmax(mx.z, r.z));
        }
        return Aabb{mn, mx};
    }
};

inline ObjectBox rotate_y(float degrees, ObjectBox object) {  // object.rs:477-484
    float radians = degrees * 3.14159265358979323846f / 180.f;  // std::f32::consts::PI
    return std::make_unique<RotateY>(std::move(object), std::sin(radians), std::cos(radians));
}

struct And : Object {  // object.rs:394-417
    ObjectBox a, b;
    And(ObjectBox a_, ObjectBox b_) : a(std::move(a_)), b(std::move(b_)) {}
    bool hit(const Ray& ray, float t_start, float t_end, const PathRng& rng, HitRecord& rec) const override {
        HitRecord hit0, hit1;
        bool h0 = a->hit(ray, t_start, t_end, rng, hit0);
        if (h0) t_end = hit0.t;
        bool h1 = b->hit(ray, t_start, t_end, rng, hit1);
        if (h1) { rec = hit1; return true; }  // hit1.or(hit0)
        if (h0) { rec = hit0; return true; }
        return false;
    }
    Aabb bounding_box(float e0, float e1) const override { return a->bounding_box(e0, e1).merge(b->bounding_box(e0, e1)); }
};

inline ObjectBox rect_prism(Vec3 p0, Vec3 p1, const Material& material) {  // object.rs:420-473
    auto R = [&](int axis, float a0, float a1, float b0, float b1, float k) -> ObjectBox {
        return std::make_unique<Rect>(axis, a0, a1, b0, b1, k, material);
    };
    auto F = [](ObjectBox o) -> ObjectBox { return std::make_unique<FlipNormals>(std::move(o)); };
    auto A = [](ObjectBox a, ObjectBox b) -> ObjectBox { return std::make_unique<And>(std::move(a), std::move(b)); };
    return A(A(R(2, p0.x, p1.x, p0.y, p1.y, p1.z),
               A(R(1, p0.x, p1.x, p0.z, p1.z, p1.y), R(0, p0.y, p1.y, p0.z, p1.z, p1.x))),
             A(F(R(2, p0.x, p1.x, p0.y, p1.y, p0.z)),
               A(F(R(1, p0.x, p1.x, p0.z, p1.z, p0.y)), F(R(0, p0.y, p1.y, p0.z, p1.z, p0.x)))));
}

struct LinearMove : Object {  // object.rs:489-528
    ObjectBox object;
    Vec3 motion;
    LinearMove(ObjectBox o, Vec3 m) : object(std::move(o)), motion(m) {}
    bool hit(const Ray& ray, float t_start, float t_end, const PathRng& rng, HitRecord& rec) const override {
        Ray r = ray;
        r.origin = ray.origin - ray.time * motion;
        return object->hit(r, t_start, t_end, rng, rec);  // result NOT moved back (reference quirk)
    }
    Aabb bounding_box(float e0, float e1) const override {
        Aabb bb = object->bounding_box(e0, e1);
        Aabb s{bb.min + e0 * motion, bb.max + e0 * motion};
        Aabb e{bb.min + e1 * motion, bb.max + e1 * motion};
        return s.merge(e);
    }
};

struct ConstantMedium : Object {  // object.rs:533-580
    ObjectBox boundary;
    float density;
    Material material;
    uint32_t medium_id;  // RNG stream id (contract): order of construction within the scene
    ConstantMedium(ObjectBox b, float d, Material m, uint32_t id)
        : boundary(std::move(b)), density(d), material(std::move(m)), medium_id(id) {}
    bool hit(const Ray& ray, float t_start, float t_end, const PathRng& rng, HitRecord& rec) const override {
        if (rng.counters) rng.counters->medium_evals++;
        HitRecord hit1, hit2;
        if (boundary->hit(ray, F32_MIN, F32_MAX, rng, hit1)) {
            if (boundary->hit(ray, hit1.t + 0.0001f, F32_MAX, rng, hit2)) {
                hit1.t = rmax(hit1.t, t_start);
                hit2.t = rmin(hit2.t, t_end);
                if (hit1.t >= hit2.t) return false;
                float distance_inside = (hit2.t - hit1.t) * length(ray.direction);
                uint32_t w[4];
                rng.block(PURPOSE_MEDIUM0 + medium_id, 0, w);
                rng.count(1);
                float hit_distance = -(1.f / density) * log_f32(u32_to_unit_f32(w[0]));
                if (hit_distance < distance_inside) {
                    float t = hit1.t + hit_distance / length(ray.direction);
                    rec.t = t;
                    rec.p = ray.point_at_parameter(t);
                    rec.normal = Vec3(1.f, 0.f, 0.f);
                    rec.material = &material;
                    return true;
                }
            }
        }
        return false;
    }
    Aabb bounding_box(float e0, float e1) const override { return boundary->bounding_box(e0, e1); }
};

// ---------------------------------------------------------------------------------------------
// Bvh — src/bvh.rs
// ---------------------------------------------------------------------------------------------
struct Bvh : Object {
    Aabb bbox;
    size_t size = 0;
    std::unique_ptr<Bvh> left, right;  // BvhContents::Node
    ObjectBox leaf;                    // BvhContents::Leaf

    // bvh.rs:22-81.  One deviation, part of the contract: the reference's sort_unstable_by is
    // pdqsort, whose order among equal keys is unspecified and std-version specific; here equal
    // keys keep their input order (std::stable_sort).  For <= 20 elements pdqsort is an insertion
    // sort and agrees.
    static std::unique_ptr<Bvh> build(std::vector<ObjectBox> objs, float e0, float e1) {
        if (objs.empty()) throw std::runtime_error("Can't create a BVH from zero objects.");  // bvh.rs:60
        auto axis_range = [&](int axis) {
            float start = F32_MAX, end = F32_MIN;
            for (auto& o : objs) {
                Aabb bb = o->bounding_box(e0, e1);
                float mn = rmin(bb.min[axis], bb.max[axis]);
                float mx = rmax(bb.min[axis], bb.max[axis]);
                start = rmin(start, mn);
                end = rmax(end, mx);
            }
            return end - start;
        };
        float ranges[3] = {axis_range(0), axis_range(1), axis_range(2)};
        for (float r : ranges)
            if (std::isnan(r)) throw std::runtime_error("NaN extent in Bvh::new (partial_cmp().unwrap())");
        int axis = 0;  // descending sort, first of equal maxima wins
        if (ranges[1] > ranges[axis]) axis = 1;
        if (ranges[2] > ranges[axis]) axis = 2;

        std::vector<std::pair<float, size_t>> keys(objs.size());
        for (size_t i = 0; i < objs.size(); ++i) {
            Aabb bb = objs[i]->bounding_box(e0, e1);
            keys[i] = {bb.min[axis] + bb.max[axis], i};  // centroid*2
            if (std::isnan(keys[i].first)) throw std::runtime_error("NaN centroid in Bvh::new");
        }
        std::stable_sort(keys.begin(), keys.end(),
                         [](const std::pair<float, size_t>& a, const std::pair<float, size_t>& b) { return a.first < b.first; });
        std::vector<ObjectBox> sorted;
        sorted.reserve(objs.size());
        for (auto& kv : keys) sorted.push_back(std::move(objs[kv.second]));

        auto node = std::make_unique<Bvh>();
        if (sorted.size() == 1) {
            node->bbox = sorted[0]->bounding_box(e0, e1);
            node->size = 1;
            node->leaf = std::move(sorted[0]);
            return node;
        }
        size_t half = sorted.size() / 2;
        std::vector<ObjectBox> right_objs;
        for (size_t i = half; i < sorted.size(); ++i) right_objs.push_back(std::move(sorted[i]));
        sorted.resize(half);
        node->right = build(std::move(right_objs), e0, e1);  // right built first (bvh.rs:68-72)
        node->left = build(std::move(sorted), e0, e1);
        node->bbox = node->left->bbox.merge(node->right->bbox);
        node->size = node->left->size + node->right->size;
        return node;
    }

    bool hit(const Ray& ray, float t_start, float t_end, const PathRng& rng, HitRecord& rec) const override {  // bvh.rs:85-120
        if (rng.counters) rng.counters->node_tests++;
        if (!bbox.hit(ray, t_start, t_end)) return false;
        if (leaf) return leaf->hit(ray, t_start, t_end, rng, rec);
        HitRecord hl, hr;
        bool l = left->hit(ray, t_start, t_end, rng, hl);
        if (l) t_end = hl.t;
        bool r = right->hit(ray, t_start, t_end, rng, hr);
        if (l && r) { rec = (hl.t < hr.t) ? hl : hr; return true; }
        if (l) { rec = hl; return true; }
        if (r) { rec = hr; return true; }
        return false;
    }
    Aabb bounding_box(float, float) const override { return bbox; }
    size_t node_count() const { return leaf ? 1 : 1 + left->node_count() + right->node_count(); }
    size_t depth() const { return leaf ? 1 : 1 + std::max(left->depth(), right->depth()); }
};

// ---------------------------------------------------------------------------------------------
// Camera — src/camera.rs
// ---------------------------------------------------------------------------------------------
struct Camera {
    Vec3 origin, lower_left_corner, horizontal, vertical, u, v;
    float lens_radius = 0.f, exposure_start = 0.f, exposure_end = 1.f;

    static Camera look(Vec3 look_from, Vec3 look_at, Vec3 up, float fov, float aspect, float aperture,
                       float focus_dist, float e0, float e1) {  // camera.rs:18-50
        Camera c;
        c.lens_radius = aperture / 2.f;
        float theta = fov * 3.14159265358979323846f / 180.f;
        float half_height = std::tan(theta / 2.f);
        float half_width = aspect * half_height;
        c.origin = look_from;
        Vec3 w = into_unit(look_from - look_at);
        c.u = into_unit(cross(up, w));
        c.v = cross(w, c.u);
        c.lower_left_corner = c.origin - half_width * focus_dist * c.u - half_height * focus_dist * c.v - focus_dist * w;
        c.horizontal = 2.f * half_width * focus_dist * c.u;
        c.vertical = 2.f * half_height * focus_dist * c.v;
        c.exposure_start = e0;
        c.exposure_end = e1;
        return c;
    }

    Ray get_ray(float s, float t, const PathRng& rng, uint32_t time_word0, uint32_t time_word1) const {  // camera.rs:52-63
        Vec3 rd = lens_radius * in_unit_disc(rng);
        Vec3 offset = rd.x * u + rd.y * v;
        // rng.gen_range(start, end): rand 0.6.5 UniformFloat::sample_single
        if (!(exposure_start < exposure_end))
            throw std::runtime_error("Uniform::sample_single called with low >= high");
        float scale = exposure_end - exposure_start;
        float off = exposure_start - scale;
        float time;
        for (uint32_t k = 0;; ++k) {
            uint32_t word;
            if (k == 0) word = time_word0;
            else if (k == 1) word = time_word1;
            else {
                uint32_t w[4];
                rng.block(PURPOSE_CAMERA, 1 + (k - 2) / 4, w);
                word = w[(k - 2) % 4];
            }
            rng.count(1);
            float res = u32_to_f32_1_2(word) * scale + off;
            if (res < exposure_end) { time = res; break; }
        }
        Ray r;
        r.origin = origin + offset;
        r.direction = lower_left_corner + s * horizontal + t * vertical - origin - offset;
        r.time = time;
        return r;
    }
};

// ---------------------------------------------------------------------------------------------
// World / hit_top / color — src/lib.rs:23-101
// ---------------------------------------------------------------------------------------------
enum class Background { Black = 0, SkyGradient = 1 };

struct World {
    std::vector<ObjectBox> list;  // impl World for [Box<dyn Object>]   lib.rs:33-49
    std::unique_ptr<Bvh> bvh;     // impl World for bvh::Bvh            lib.rs:51-55

    bool hit_top(const Ray& ray, const PathRng& rng, HitRecord& rec) const {
        if (bvh) return bvh->hit(ray, 0.001f, F32_MAX, rng, rec);
        const float NEAR = 0.001f;
        float nearest = F32_MAX;
        bool any = false;
        for (auto& obj : list) {
            HitRecord r;
            if (obj->hit(ray, NEAR, nearest, rng, r)) {
                nearest = r.t;
                rec = r;
                any = true;
            }
        }
        return any;
    }
};

// lib.rs:60-101.  HEAD returns black when the ray escapes (lib.rs:100); the book-1 image in the
// README was rendered with the book's sky gradient, offered here as Background::
// SkyGradient: strength * ((1-t)*(1,1,1) + t*(0.5,0.7,1.0)), t = 0.5*(unit(dir).y + 1).
inline Vec3 color(const World& world, Ray ray, PathRng& rng, Background bg) {
    Vec3 accum;
    Vec3 strength = splat(1.f);
    uint32_t bounces = 0;
    HitRecord hit;
    uint64_t segs = 0;
    for (;;) {
        rng.bounce = bounces;
        ++segs;
        if (!world.hit_top(ray, rng, hit)) break;
        accum = accum + strength * hit.material->emitted(hit.p);
        Ray new_ray;
        Vec3 attenuation;
        if (hit.material->scatter(ray, hit, rng, new_ray, attenuation)) {
            ray = new_ray;
            strength = strength * attenuation;
        } else {
            if (rng.counters) { rng.counters->segments += segs; rng.counters->max_segments = std::max(rng.counters->max_segments, segs); }
            return accum;
        }
        if (bounces == 50) {
            if (rng.counters) { rng.counters->segments += segs; rng.counters->max_segments = std::max(rng.counters->max_segments, segs); }
            return accum;
        }
        bounces += 1;
    }
    if (rng.counters) { rng.counters->segments += segs; rng.counters->max_segments = std::max(rng.counters->max_segments, segs); }
    if (bg == Background::SkyGradient) {
        Vec3 unit_direction = into_unit(ray.direction);
        float t = 0.5f * (unit_direction.y + 1.0f);
        return strength * ((1.0f - t) * Vec3(1.f, 1.f, 1.f) + t * Vec3(0.5f, 0.7f, 1.0f));
    }
    return Vec3();
}

// One pixel-sample of cast()/par_cast(): lib.rs:387-393 / :366-372.
inline Vec3 sample_color(const World& world, const Camera& camera, Background bg, uint32_t nx, uint32_t ny,
                         uint32_t x, uint32_t y, uint32_t s, uint64_t seed, Counters* counters) {
    PathRng rng(seed, y * nx + x, s, counters);
    uint32_t w[4];
    rng.block(PURPOSE_CAMERA, 0, w);
    rng.count(2);
    float u = (static_cast<float>(x) + u32_to_unit_f32(w[0])) / static_cast<float>(nx);
    float v = (static_cast<float>(y) + u32_to_unit_f32(w[1])) / static_cast<float>(ny);
    Ray r = camera.get_ray(u, v, rng, w[2], w[3]);
    if (counters) counters->samples++;
    return color(world, r, rng, bg);
}

// lib.rs:350-352
inline int to_u8(float x) {
    int v = f32_as_i32(255.99f * x);
    return std::min(std::max(v, 0), 255);
}

}  // namespace oracle
