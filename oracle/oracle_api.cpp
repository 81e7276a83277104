// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.hpp for the rules).  PARITY UNPINNED.
//
// C entry points (ctypes-friendly) over the CPU restatement: scene construction by name, the
// cast()/par_cast() render loop (src/lib.rs:324-397), print_ppm's quantisation (lib.rs:344-361),
// work counters, and single-function probes used by the known-answer tests.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <thread>

#include "oracle_scenes.hpp"

using namespace oracle;

namespace {
thread_local std::string g_err;
int fail(const std::exception& e) { g_err = e.what(); return 1; }
}  // namespace

extern "C" {

const char* oracle_last_error() { return g_err.c_str(); }

void* oracle_scene_build(const char* name, uint32_t nx, uint32_t ny, uint64_t scene_seed, int top_level_bvh) {
    try {
        return scenes::build(name, nx, ny, scene_seed, top_level_bvh != 0).release();
    } catch (const std::exception& e) {
        fail(e);
        return nullptr;
    }
}
void oracle_scene_free(void* s) { delete static_cast<Scene*>(s); }

// background: -1 = the scene's own, 0 = Black, 1 = SkyGradient
// out_rgb: (row_end-row_begin) * nx * 3 floats, row 0 = TOP scanline (lib.rs:326-330), linear,
//          divided by ns (lib.rs:374).  May be null.
// out_samples: (row_end-row_begin) * nx * ns * 3 floats, per-sample radiance before the sum. May be null.
// counters: 8 x u64 {samples, segments, node_tests, sphere_tests, rect_tests, medium_evals, draws,
//           max_segments}. May be null.
// nthreads: 1 = cast()'s sequential scan order; >1 = one task per scanline like par_compute
//           (lib.rs:327).  The result is identical either way (the RNG is keyed by pixel/sample).
int oracle_render(void* scene, uint32_t nx, uint32_t ny, uint32_t ns, uint64_t seed, int background,
                  uint32_t row_begin, uint32_t row_end, int nthreads, float* out_rgb, float* out_samples,
                  uint64_t* counters) {
    try {
        const Scene& sc = *static_cast<const Scene*>(scene);
        if (row_end > ny || row_begin > row_end) throw std::runtime_error("bad row range");
        Background bg = background < 0 ? sc.background : static_cast<Background>(background);
        std::atomic<uint32_t> next_row{row_begin};
        if (nthreads < 1) nthreads = 1;
        std::vector<Counters> per_thread(nthreads);
        std::string thread_err;
        std::atomic<bool> failed{false};
        auto work = [&](int tid) {
            Counters* cnt = counters ? &per_thread[tid] : nullptr;
            try {
                for (;;) {
                    uint32_t row = next_row.fetch_add(1);
                    if (row >= row_end) break;
                    uint32_t y = ny - 1 - row;  // (0..ny).rev()
                    for (uint32_t x = 0; x < nx; ++x) {
                        Vec3 col;  // iter.fold(Vec3::default(), Add)   vec3.rs:195-203
                        for (uint32_t s = 0; s < ns; ++s) {
                            Vec3 c = sample_color(sc.world, sc.camera, bg, nx, ny, x, y, s, seed, cnt);
                            if (out_samples) {
                                float* o = out_samples + ((static_cast<size_t>(row - row_begin) * nx + x) * ns + s) * 3;
                                o[0] = c.x; o[1] = c.y; o[2] = c.z;
                            }
                            col = col + c;
                        }
                        col = col / static_cast<float>(ns);
                        if (out_rgb) {
                            float* o = out_rgb + (static_cast<size_t>(row - row_begin) * nx + x) * 3;
                            o[0] = col.x; o[1] = col.y; o[2] = col.z;
                        }
                    }
                }
            } catch (const std::exception& e) {
                if (!failed.exchange(true)) thread_err = e.what();
            }
        };
        if (nthreads == 1) {
            work(0);
        } else {
            std::vector<std::thread> ts;
            for (int t = 0; t < nthreads; ++t) ts.emplace_back(work, t);
            for (auto& t : ts) t.join();
        }
        if (failed) throw std::runtime_error(thread_err);
        if (counters) {
            Counters total;
            for (auto& c : per_thread) total.add(c);
            counters[0] = total.samples; counters[1] = total.segments; counters[2] = total.node_tests;
            counters[3] = total.sphere_tests; counters[4] = total.rect_tests; counters[5] = total.medium_evals;
            counters[6] = total.draws; counters[7] = total.max_segments;
        }
        return 0;
    } catch (const std::exception& e) {
        return fail(e);
    }
}

// lib.rs:344-361: sqrt gamma then to_u8.  in: n floats (linear), out: n bytes-as-int32.
void oracle_ppm_quantise(const float* linear, size_t n, int32_t* out) {
    for (size_t i = 0; i < n; ++i) out[i] = to_u8(std::sqrt(linear[i]));
}

// Scene introspection for tests.
int oracle_scene_info(void* scene, uint64_t* out /* [top_is_bvh, n_top_objects, bvh_nodes, bvh_depth, n_media] */) {
    const Scene& sc = *static_cast<const Scene*>(scene);
    out[0] = sc.world.bvh ? 1 : 0;
    out[1] = sc.world.bvh ? sc.world.bvh->size : sc.world.list.size();
    out[2] = sc.world.bvh ? sc.world.bvh->node_count() : 0;
    out[3] = sc.world.bvh ? sc.world.bvh->depth() : 0;
    out[4] = sc.n_media;
    return 0;
}
void oracle_scene_camera(void* scene, float* out /* 21 floats: origin,llc,hor,ver,u,v,lens,e0,e1 */) {
    const Camera& c = static_cast<const Scene*>(scene)->camera;
    const Vec3* vs[6] = {&c.origin, &c.lower_left_corner, &c.horizontal, &c.vertical, &c.u, &c.v};
    for (int i = 0; i < 6; ++i) { out[3 * i] = vs[i]->x; out[3 * i + 1] = vs[i]->y; out[3 * i + 2] = vs[i]->z; }
    out[18] = c.lens_radius; out[19] = c.exposure_start; out[20] = c.exposure_end;
}
void oracle_scene_perlin(void* scene, float* vecs /*768*/, uint8_t* perms /*768: x,y,z*/) {
    const PerlinTables& tb = *static_cast<const Scene*>(scene)->perlin;
    for (int i = 0; i < 256; ++i) { vecs[3 * i] = tb.vecs[i].x; vecs[3 * i + 1] = tb.vecs[i].y; vecs[3 * i + 2] = tb.vecs[i].z; }
    std::memcpy(perms, tb.perm_x, 256); std::memcpy(perms + 256, tb.perm_y, 256); std::memcpy(perms + 512, tb.perm_z, 256);
}

// ------------------------------- single-function probes (KATs) -------------------------------
void oracle_philox4x32_10(const uint32_t* key, const uint32_t* ctr, uint32_t* out) { philox4x32_10(key, ctr, out); }
float oracle_u32_to_unit_f32(uint32_t w) { return u32_to_unit_f32(w); }
float oracle_u32_to_f32_1_2(uint32_t w) { return u32_to_f32_1_2(w); }
float oracle_log_f32(float x) { return log_f32(x); }
float oracle_sin_f32(float x) { return sin_f32(x); }
float oracle_pow5_f32(float x) { return pow5_f32(x); }
float oracle_schlick(float c, float ri) { return schlick(c, ri); }
int oracle_to_u8(float x) { return to_u8(x); }
float oracle_dot(const float* a, const float* b) { return dot(Vec3(a[0], a[1], a[2]), Vec3(b[0], b[1], b[2])); }
void oracle_cross(const float* a, const float* b, float* o) {
    Vec3 r = cross(Vec3(a[0], a[1], a[2]), Vec3(b[0], b[1], b[2]));
    o[0] = r.x; o[1] = r.y; o[2] = r.z;
}
void oracle_into_unit(const float* a, float* o) {
    Vec3 r = into_unit(Vec3(a[0], a[1], a[2]));
    o[0] = r.x; o[1] = r.y; o[2] = r.z;
}
void oracle_reflect(const float* v, const float* n, float* o) {
    Vec3 r = reflect(Vec3(v[0], v[1], v[2]), Vec3(n[0], n[1], n[2]));
    o[0] = r.x; o[1] = r.y; o[2] = r.z;
}
int oracle_refract(const float* v, const float* n, float ni_over_nt, float* o) {
    Vec3 r;
    bool ok = refract(Vec3(v[0], v[1], v[2]), Vec3(n[0], n[1], n[2]), ni_over_nt, r);
    o[0] = r.x; o[1] = r.y; o[2] = r.z;
    return ok ? 1 : 0;
}
// ray = {ox,oy,oz,dx,dy,dz,time}
static Ray mk_ray(const float* r) {
    Ray ray;
    ray.origin = Vec3(r[0], r[1], r[2]);
    ray.direction = Vec3(r[3], r[4], r[5]);
    ray.time = r[6];
    return ray;
}
int oracle_aabb_hit(const float* mn, const float* mx, const float* ray, float t0, float t1) {
    Aabb b{Vec3(mn[0], mn[1], mn[2]), Vec3(mx[0], mx[1], mx[2])};
    return b.hit(mk_ray(ray), t0, t1) ? 1 : 0;
}
static int put_hit(bool ok, const HitRecord& h, float* out) {
    if (ok) { out[0] = h.t; out[1] = h.p.x; out[2] = h.p.y; out[3] = h.p.z; out[4] = h.normal.x; out[5] = h.normal.y; out[6] = h.normal.z; }
    return ok ? 1 : 0;
}
int oracle_sphere_hit(float radius, const float* ray, float t0, float t1, float* out7) {
    Sphere s(radius, Material::dielectric(1.5f));
    PathRng rng(0, 0, 0, nullptr);
    HitRecord h;
    return put_hit(s.hit(mk_ray(ray), t0, t1, rng, h), h, out7);
}
int oracle_rect_hit(int axis, float r0s, float r0e, float r1s, float r1e, float k, int flip, const float* ray,
                    float t0, float t1, float* out7) {
    ObjectBox o = std::make_unique<Rect>(axis, r0s, r0e, r1s, r1e, k, Material::dielectric(1.5f));
    if (flip) o = std::make_unique<FlipNormals>(std::move(o));
    PathRng rng(0, 0, 0, nullptr);
    HitRecord h;
    return put_hit(o->hit(mk_ray(ray), t0, t1, rng, h), h, out7);
}
// wrapper: 0 Translate(v), 1 Scale(v), 2 RotateY(v[0] degrees), 3 LinearMove(v), 4 FlipNormals — around a sphere
int oracle_wrapped_sphere_hit(int wrapper, const float* v, float radius, const float* ray, float t0, float t1, float* out7) {
    ObjectBox o = std::make_unique<Sphere>(radius, Material::dielectric(1.5f));
    Vec3 vv(v[0], v[1], v[2]);
    switch (wrapper) {
        case 0: o = std::make_unique<Translate>(vv, std::move(o)); break;
        case 1: o = std::make_unique<Scale>(vv, std::move(o)); break;
        case 2: o = rotate_y(v[0], std::move(o)); break;
        case 3: o = std::make_unique<LinearMove>(std::move(o), vv); break;
        default: o = std::make_unique<FlipNormals>(std::move(o)); break;
    }
    PathRng rng(0, 0, 0, nullptr);
    HitRecord h;
    return put_hit(o->hit(mk_ray(ray), t0, t1, rng, h), h, out7);
}
// ConstantMedium{Sphere(radius)} with the medium draw forced: the probe replaces the Philox word
// by searching nothing — instead it reports the pieces: returns 1 and t for a given uniform u.
int oracle_medium_hit_with_u(float radius, float density, const float* ray_, float t0, float t1, float u, float* out_t) {
    Ray ray = mk_ray(ray_);
    Sphere s(radius, Material::dielectric(1.5f));
    PathRng rng(0, 0, 0, nullptr);
    HitRecord h1, h2;
    if (!s.hit(ray, F32_MIN, F32_MAX, rng, h1)) return 0;
    if (!s.hit(ray, h1.t + 0.0001f, F32_MAX, rng, h2)) return 0;
    h1.t = rmax(h1.t, t0);
    h2.t = rmin(h2.t, t1);
    if (h1.t >= h2.t) return 0;
    float distance_inside = (h2.t - h1.t) * length(ray.direction);
    float hit_distance = -(1.f / density) * log_f32(u);
    if (hit_distance < distance_inside) { *out_t = h1.t + hit_distance / length(ray.direction); return 1; }
    return 0;
}
float oracle_perlin_noise(void* scene, const float* p) {
    return perlin_noise(*static_cast<const Scene*>(scene)->perlin, Vec3(p[0], p[1], p[2]));
}
float oracle_perlin_turb(void* scene, const float* p, int depth) {
    return perlin_turb(*static_cast<const Scene*>(scene)->perlin, Vec3(p[0], p[1], p[2]), depth);
}
void oracle_checker(const float* p, const float* c0, const float* c1, float* out) {
    Texture t = tex_checker(tex_constant(Vec3(c0[0], c0[1], c0[2])), tex_constant(Vec3(c1[0], c1[1], c1[2])));
    Vec3 r = (*t)(Vec3(p[0], p[1], p[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void oracle_camera_look(const float* from, const float* at, const float* up, float fov, float aspect, float aperture,
                        float focus, float e0, float e1, float* out21) {
    Scene tmp;
    tmp.camera = Camera::look(Vec3(from[0], from[1], from[2]), Vec3(at[0], at[1], at[2]), Vec3(up[0], up[1], up[2]), fov,
                              aspect, aperture, focus, e0, e1);
    oracle_scene_camera(&tmp, out21);
}
// One camera ray for (pixel x,y counted from the bottom; sample s): {o, d, time}
int oracle_get_ray(void* scene, uint32_t nx, uint32_t ny, uint32_t x, uint32_t y, uint32_t s, uint64_t seed, float* out7) {
    try {
        const Scene& sc = *static_cast<const Scene*>(scene);
        PathRng rng(seed, y * nx + x, s, nullptr);
        uint32_t w[4];
        rng.block(PURPOSE_CAMERA, 0, w);
        float u = (static_cast<float>(x) + u32_to_unit_f32(w[0])) / static_cast<float>(nx);
        float v = (static_cast<float>(y) + u32_to_unit_f32(w[1])) / static_cast<float>(ny);
        Ray r = sc.camera.get_ray(u, v, rng, w[2], w[3]);
        out7[0] = r.origin.x; out7[1] = r.origin.y; out7[2] = r.origin.z;
        out7[3] = r.direction.x; out7[4] = r.direction.y; out7[5] = r.direction.z; out7[6] = r.time;
        return 0;
    } catch (const std::exception& e) {
        return fail(e);
    }
}
// First hit of a ray against the whole scene (hit_top): returns 1 and {t,p,n}.
int oracle_hit_top(void* scene, const float* ray, float* out7) {
    const Scene& sc = *static_cast<const Scene*>(scene);
    PathRng rng(0, 0, 0, nullptr);
    HitRecord h;
    return put_hit(sc.world.hit_top(mk_ray(ray), rng, h), h, out7);
}
void oracle_smallrng_u32(uint64_t seed, uint32_t n, uint32_t* out) {
    SmallRng r = SmallRng::seed_from_u64(seed);
    for (uint32_t i = 0; i < n; ++i) out[i] = r.next_u32();
}

}  // extern "C"

#ifdef ORACLE_CLI
// oracle_cli <scene> <nx> <ny> <ns> [seed] [threads] [top_bvh] : renders and prints a P3 PPM
// exactly as print_ppm does (lib.rs:344-361); timing to stderr like main.rs:353.
int main(int argc, char** argv) {
    if (argc < 5) { std::fprintf(stderr, "usage: %s scene nx ny ns [seed] [threads] [top_bvh]\n", argv[0]); return 2; }
    uint32_t nx = std::atoi(argv[2]), ny = std::atoi(argv[3]), ns = std::atoi(argv[4]);
    uint64_t seed = argc > 5 ? std::strtoull(argv[5], nullptr, 0) : 0xDEADBEEFull;
    int threads = argc > 6 ? std::atoi(argv[6]) : 1;
    int top_bvh = argc > 7 ? std::atoi(argv[7]) : 1;
    void* sc = oracle_scene_build(argv[1], nx, ny, 0xDEADBEEFull, top_bvh);
    if (!sc) { std::fprintf(stderr, "%s\n", oracle_last_error()); return 1; }
    std::vector<float> img(static_cast<size_t>(nx) * ny * 3);
    uint64_t cnt[8];
    auto t0 = std::chrono::steady_clock::now();
    if (oracle_render(sc, nx, ny, ns, seed, -1, 0, ny, threads, img.data(), nullptr, cnt)) { std::fprintf(stderr, "%s\n", oracle_last_error()); return 1; }
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::fprintf(stderr, "Took %.3fs wall time (%d threads): %.3f Msamples/s; per sample: %.2f segments, %.2f node tests, %.2f sphere, %.2f rect, %.2f medium, %.2f draws\n",
                 dt, threads, cnt[0] / dt / 1e6, double(cnt[1]) / cnt[0], double(cnt[2]) / cnt[0], double(cnt[3]) / cnt[0],
                 double(cnt[4]) / cnt[0], double(cnt[5]) / cnt[0], double(cnt[6]) / cnt[0]);
    std::printf("P3\n%u %u\n255\n", nx, ny);
    for (size_t i = 0; i < img.size(); i += 3)
        std::printf("%d %d %d\n", to_u8(std::sqrt(img[i])), to_u8(std::sqrt(img[i + 1])), to_u8(std::sqrt(img[i + 2])));
    oracle_scene_free(sc);
    return 0;
}
#endif
