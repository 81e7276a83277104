// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle_math.hpp for the rules).  PARITY UNPINNED.
//
// Scene builders restated from the reference (src/lib.rs:103-193, src/lib.rs:237-319 [commented
// out at HEAD], src/main.rs:10-319, benches/scene.rs:13-30) plus the scene-generation RNG
// (rand 0.6.5 SmallRng = rand_pcg 0.1.2 Pcg64Mcg; un-vendored, restated from its published
// algorithm — SURVEY App. C; UNVERIFIED against the crate because it is not in /root/reference).
#pragma once
#include "oracle_core.hpp"

namespace oracle {

struct SmallRng {  // rand_pcg::Mcg128Xsl64
    unsigned __int128 state;

    static SmallRng seed_from_u64(uint64_t s) {  // rand_core::SeedableRng::seed_from_u64 (PCG32 fill)
        const uint64_t MUL = 6364136223846793005ull, INC = 11634580027462260723ull;
        uint8_t seed[16];
        for (int chunk = 0; chunk < 4; ++chunk) {
            s = s * MUL + INC;
            uint32_t xorshifted = static_cast<uint32_t>(((s >> 18) ^ s) >> 27);
            uint32_t rot = static_cast<uint32_t>(s >> 59);
            uint32_t x = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
            for (int b = 0; b < 4; ++b) seed[chunk * 4 + b] = static_cast<uint8_t>(x >> (8 * b));
        }
        unsigned __int128 st = 0;
        for (int b = 15; b >= 0; --b) st = (st << 8) | seed[b];
        SmallRng r;
        r.state = st | 1;
        return r;
    }
    uint64_t next_u64() {
        const unsigned __int128 MULT =
            (static_cast<unsigned __int128>(0x2360ED051FC65DA4ull) << 64) | 0x4385DF649FCCF645ull;
        state *= MULT;
        uint32_t rot = static_cast<uint32_t>(state >> 122);
        uint64_t xsl = static_cast<uint64_t>(state >> 64) ^ static_cast<uint64_t>(state);
        return (xsl >> rot) | (xsl << ((64 - rot) & 63));
    }
    uint32_t next_u32() { return static_cast<uint32_t>(next_u64()); }
    float gen_f32() { return u32_to_unit_f32(next_u32()); }
    Vec3 gen_vec3() {  // vec3.rs:209-214: x, y, z in order
        float a = gen_f32(), b = gen_f32(), c = gen_f32();
        return Vec3(a, b, c);
    }
    float gen_range_f32(float low, float high) {  // UniformFloat::sample_single
        float scale = high - low, offset = low - scale;
        for (;;) {
            float res = u32_to_f32_1_2(next_u32()) * scale + offset;
            if (res < high) return res;
        }
    }
    uint64_t gen_range_usize(uint64_t low, uint64_t high) {  // UniformInt<usize>::sample_single
        uint64_t range = high - low;
        uint64_t zone = (range << __builtin_clzll(range)) - 1;
        for (;;) {
            unsigned __int128 m = static_cast<unsigned __int128>(next_u64()) * range;
            if (static_cast<uint64_t>(m) <= zone) return low + static_cast<uint64_t>(m >> 64);
        }
    }
};

// perlin.rs:5-29, seeded.  Order: VECS, PERM_X, PERM_Y, PERM_Z from one generator.
inline std::shared_ptr<const PerlinTables> make_perlin_tables(uint64_t scene_seed) {
    SmallRng rng = SmallRng::seed_from_u64(scene_seed ^ 0x5045524C494Eull /* "PERLIN" */);
    auto tb = std::make_shared<PerlinTables>();
    for (int i = 0; i < 256; ++i) {  // generate_vecs: Vec3::in_unit_sphere on a sequential rng
        for (;;) {
            Vec3 v = 2.f * rng.gen_vec3() - splat(1.f);
            if (dot(v, v) < 1.f) { tb->vecs[i] = v; break; }
        }
    }
    auto perm = [&](uint8_t* p) {  // generate_perm
        for (int i = 0; i < 256; ++i) p[i] = static_cast<uint8_t>(i);
        for (int i = 255; i >= 1; --i) std::swap(p[i], p[rng.gen_range_usize(0, static_cast<uint64_t>(i))]);
    };
    perm(tb->perm_x);
    perm(tb->perm_y);
    perm(tb->perm_z);
    return tb;
}

struct Scene {
    World world;
    Camera camera;
    Background background = Background::Black;
    std::shared_ptr<const PerlinTables> perlin;
    uint32_t n_media = 0;
};

namespace scenes {

inline ObjectBox sphere(float r, const Material& m) { return std::make_unique<Sphere>(r, m); }
inline ObjectBox translate(Vec3 off, ObjectBox o) { return std::make_unique<Translate>(off, std::move(o)); }
inline ObjectBox flip(ObjectBox o) { return std::make_unique<FlipNormals>(std::move(o)); }
inline ObjectBox rect(int axis, float a0, float a1, float b0, float b1, float k, const Material& m) {
    return std::make_unique<Rect>(axis, a0, a1, b0, b1, k, m);
}
inline Material diffuse_color(Vec3 c) { return Material::lambertian(tex_constant(c)); }

inline std::vector<ObjectBox> cornell_box() {  // lib.rs:103-166
    Material red = diffuse_color(Vec3(0.65f, 0.05f, 0.05f));
    Material white = diffuse_color(splat(0.73f));
    Material green = diffuse_color(Vec3(0.12f, 0.45f, 0.15f));
    Material light = Material::diffuse_light(tex_constant(splat(1.f)), 15.f);
    std::vector<ObjectBox> v;
    v.push_back(rect(1, 213.f, 343.f, 227.f, 332.f, 554.f, light));
    v.push_back(rect(1, 0.f, 555.f, 0.f, 555.f, 0.f, white));          // floor
    v.push_back(flip(rect(2, 0.f, 555.f, 0.f, 555.f, 555.f, white)));  // rear wall
    v.push_back(flip(rect(1, 0.f, 555.f, 0.f, 555.f, 555.f, white)));  // ceiling
    v.push_back(rect(0, 0.f, 555.f, 0.f, 555.f, 0.f, red));            // right wall
    v.push_back(flip(rect(0, 0.f, 555.f, 0.f, 555.f, 555.f, green)));  // left wall
    return v;
}

inline std::vector<ObjectBox> cornell_box_with_boxes() {  // lib.rs:168-193
    std::vector<ObjectBox> scene = cornell_box();
    Material white = diffuse_color(splat(0.73f));
    scene.push_back(translate(Vec3(130.f, 0.f, 65.f),
                              rotate_y(-18.f, rect_prism(Vec3(0.f, 0.f, 0.f), Vec3(165.f, 165.f, 165.f), white))));
    scene.push_back(translate(Vec3(265.f, 0.f, 295.f),
                              rotate_y(15.f, rect_prism(Vec3(0.f, 0.f, 0.f), Vec3(165.f, 330.f, 165.f), white))));
    return scene;
}

inline Camera cornell_camera(uint32_t nx, uint32_t ny) {  // main.rs:12-27 (same in :34-49, :71-86, :116-131)
    return Camera::look(Vec3(278.f, 278.f, -800.f), Vec3(278.f, 278.f, 0.f), Vec3(0.f, 1.f, 0.f), 40.f,
                        static_cast<float>(nx) / static_cast<float>(ny), 0.0f, 10.f, 0.f, 1.f);
}
inline Camera book1_camera(uint32_t nx, uint32_t ny) {  // benches/scene.rs:16-30
    return Camera::look(Vec3(13.f, 2.f, 3.f), Vec3(0.f, 0.f, 0.f), Vec3(0.f, 1.f, 0.f), 20.f,
                        static_cast<float>(nx) / static_cast<float>(ny), 0.1f, 10.f, 0.f, 1.f);
}

// lib.rs:237-319 (commented out at HEAD, enum-style API) re-expressed in the live API, SURVEY
// App. B.1.  head_variant=false: canonical book-1 materials (what img/demo-scene.jpg shows);
// head_variant=true: the commented block's own materials (Perlin ground, moving Lambertians,
// Perlin light).
inline void random_scene(Scene& sc, SmallRng& rng, bool head_variant) {
    auto& world = sc.world.list;
    Material ground = head_variant ? Material::lambertian(tex_perlin(sc.perlin, 4.f)) : diffuse_color(splat(0.5f));
    world.push_back(translate(Vec3(0.f, -1000.f, 0.f), sphere(1000.f, ground)));
    for (int a = -11; a < 11; ++a) {
        for (int b = -11; b < 11; ++b) {
            float cx = static_cast<float>(a) + 0.9f * rng.gen_f32();
            float cz = static_cast<float>(b) + 0.9f * rng.gen_f32();
            Vec3 center(cx, 0.2f, cz);
            if (length(center - Vec3(4.f, 0.2f, 0.f)) > 0.9f) {
                float choose_mat = rng.gen_f32();
                if (choose_mat < 0.8f) {
                    Vec3 c0 = rng.gen_vec3();
                    Vec3 c1 = rng.gen_vec3();
                    Material m = diffuse_color(c0 * c1);
                    if (head_variant) {
                        Vec3 motion(0.f, rng.gen_range_f32(0.f, 0.5f), 0.f);
                        world.push_back(translate(center, std::make_unique<LinearMove>(sphere(0.2f, m), motion)));
                    } else {
                        world.push_back(translate(center, sphere(0.2f, m)));
                    }
                } else if (choose_mat < 0.95f) {
                    Vec3 albedo = 0.5f * (1.f + rng.gen_vec3());
                    float fuzz = 0.5f * rng.gen_f32();
                    world.push_back(translate(center, sphere(0.2f, Material::metal(albedo, fuzz))));
                } else {
                    world.push_back(translate(center, sphere(0.2f, Material::dielectric(1.5f))));
                }
            }
        }
    }
    world.push_back(translate(Vec3(0.f, 1.f, 0.f), sphere(1.0f, Material::dielectric(1.5f))));
    world.push_back(translate(Vec3(-4.f, 1.f, 0.f), sphere(1.0f, Material::metal(Vec3(0.7f, 0.6f, 0.5f), 0.f))));
    Material last = head_variant ? Material::diffuse_light(tex_perlin(sc.perlin, 10.f), 4.f)
                                 : diffuse_color(Vec3(0.4f, 0.2f, 0.1f));
    world.push_back(translate(Vec3(4.f, 1.f, 0.f), sphere(1.0f, last)));
}

inline void book_final_scene(Scene& sc, SmallRng& rng) {  // main.rs:161-319
    auto& world = sc.world.list;
    Material ground = diffuse_color(Vec3(0.48f, 0.83f, 0.53f));
    {
        std::vector<ObjectBox> boxes;
        for (int i = 0; i < 20; ++i)
            for (int j = 0; j < 20; ++j) {
                const float W = 100.f;
                Vec3 c0(-1000.f + static_cast<float>(i) * W, 0.f, -1000.f + static_cast<float>(j) * W);
                Vec3 c1 = c0 + Vec3(W, 100.f * (rng.gen_f32() + 0.01f), W);
                boxes.push_back(rect_prism(c0, c1, ground));
            }
        world.push_back(Bvh::build(std::move(boxes), 0.f, 1.f));
    }
    world.push_back(rect(1, 123.f, 423.f, 147.f, 412.f, 554.f, Material::diffuse_light(tex_constant(splat(1.f)), 7.f)));
    world.push_back(translate(Vec3(400.f, 400.f, 200.f),
                              std::make_unique<LinearMove>(sphere(50.f, diffuse_color(Vec3(0.7f, 0.3f, 0.1f))),
                                                           Vec3(30.f, 0.f, 0.f))));
    Material glass = Material::dielectric(1.5f);
    world.push_back(translate(Vec3(260.f, 150.f, 45.f), sphere(50.f, glass)));
    world.push_back(translate(Vec3(0.f, 150.f, 145.f), sphere(50.f, Material::metal(Vec3(0.8f, 0.8f, 0.9f), 1.f))));
    world.push_back(translate(Vec3(360.f, 150.f, 145.f), sphere(70.f, glass)));
    world.push_back(std::make_unique<ConstantMedium>(translate(Vec3(360.f, 150.f, 145.f), sphere(70.f, glass)), 0.2f,
                                                     Material::isotropic(tex_constant(Vec3(0.2f, 0.4f, 0.9f))),
                                                     sc.n_media++));
    world.push_back(std::make_unique<ConstantMedium>(sphere(5000.f, glass), 0.0001f,
                                                     Material::isotropic(tex_constant(splat(1.f))), sc.n_media++));
    world.push_back(translate(Vec3(220.f, 280.f, 300.f),
                              sphere(80.f, Material::lambertian(tex_perlin(sc.perlin, 0.05f)))));
    {
        Material white = diffuse_color(splat(0.73f));
        std::vector<ObjectBox> spheres;
        for (int i = 0; i < 1000; ++i) spheres.push_back(translate(165.f * rng.gen_vec3(), sphere(10.f, white)));
        world.push_back(translate(Vec3(-100.f, 270.f, 395.f), rotate_y(15.f, Bvh::build(std::move(spheres), 0.f, 1.f))));
    }
}

inline void motion_test(Scene& sc) {  // main.rs:33-67
    sc.world.list = cornell_box();
    sc.world.list.push_back(translate(
        Vec3(278.f, 278.f, 278.f),
        std::make_unique<LinearMove>(sphere(65.f, diffuse_color(splat(0.73f))), Vec3(0.f, 100.f, 0.f))));
}

inline void volume_test(Scene& sc) {  // main.rs:70-108
    sc.world.list = cornell_box();
    sc.world.list.push_back(translate(
        Vec3(278.f, 278.f, 278.f),
        std::make_unique<ConstantMedium>(sphere(180.f, diffuse_color(splat(0.73f))), 0.01f,
                                         Material::isotropic(tex_constant(Vec3(0.2f, 0.2f, 1.0f))), sc.n_media++)));
}

inline void simple_light_scene(Scene& sc, SmallRng& rng) {  // main.rs:111-159
    sc.world.list = cornell_box();
    for (int i = 0; i < 1000; ++i) {
        Vec3 off = 277.f + 257.f * rng.gen_vec3();
        sc.world.list.push_back(translate(off, sphere(20.f, diffuse_color(splat(0.3f)))));
    }
    sc.world.list.push_back(flip(sphere(1000.f, Material::diffuse_light(tex_constant(splat(0.1f)), 1.f))));
}

// Not in the reference: one scene that reaches every Object/Material/Texture implementor and
// nesting the crate allows but its own scenes never build (Scale, FlipNormals<Sphere>, checker,
// And of spheres, nested Translate, Bvh under Scale under RotateY, textured light ...).
inline void kitchen_sink(Scene& sc, SmallRng& rng) {
    auto& w = sc.world.list;
    w.push_back(rect(1, -10.f, 10.f, -10.f, 10.f, 0.f,
                     Material::lambertian(tex_checker(tex_constant(Vec3(0.2f, 0.3f, 0.1f)), tex_constant(splat(0.9f))))));
    w.push_back(std::make_unique<Scale>(
        Vec3(1.5f, 0.75f, 1.0f), translate(Vec3(-2.f, 1.5f, 0.f), sphere(1.f, Material::metal(Vec3(0.8f, 0.6f, 0.2f), 0.1f)))));
    w.push_back(rotate_y(30.f, translate(Vec3(2.f, 0.f, -1.f),
                                         rect_prism(Vec3(0.f, 0.f, 0.f), Vec3(1.f, 2.f, 1.f), Material::dielectric(1.5f)))));
    w.push_back(std::make_unique<And>(
        translate(Vec3(0.f, 1.f, 2.f), sphere(0.7f, Material::lambertian(tex_perlin(sc.perlin, 3.f)))),
        translate(Vec3(0.6f, 1.f, 2.3f), sphere(0.5f, diffuse_color(Vec3(0.7f, 0.1f, 0.1f))))));
    w.push_back(flip(sphere(40.f, Material::diffuse_light(tex_constant(Vec3(0.6f, 0.7f, 0.9f)), 1.f))));
    w.push_back(translate(
        Vec3(-1.f, 0.5f, 3.f),
        std::make_unique<LinearMove>(
            sphere(0.5f, Material::lambertian(tex_checker(tex_perlin(sc.perlin, 5.f), tex_constant(Vec3(0.1f, 0.1f, 0.8f))))),
            Vec3(0.f, 0.5f, 0.f))));
    w.push_back(std::make_unique<ConstantMedium>(translate(Vec3(3.f, 1.f, 2.f), sphere(1.f, Material::dielectric(1.5f))),
                                                 0.8f, Material::isotropic(tex_constant(splat(0.9f))), sc.n_media++));
    w.push_back(translate(Vec3(1.f, 3.f, 0.f),
                          flip(rect(1, -0.5f, 0.5f, -0.5f, 0.5f, 0.f, Material::diffuse_light(tex_perlin(sc.perlin, 2.f), 8.f)))));
    w.push_back(translate(Vec3(0.f, 0.2f, 0.f),
                          translate(Vec3(-3.f, 0.f, 3.f), sphere(0.2f, Material::metal(Vec3(0.9f, 0.9f, 0.9f), 0.9f)))));
    {
        std::vector<ObjectBox> small;
        for (int i = 0; i < 8; ++i) {
            Vec3 off = 2.f * rng.gen_vec3();
            Vec3 col = rng.gen_vec3();
            small.push_back(translate(off, sphere(0.25f, diffuse_color(col))));
        }
        w.push_back(translate(Vec3(-4.f, 0.f, -2.f),
                              rotate_y(-40.f, std::make_unique<Scale>(Vec3(1.f, 2.f, 1.f), Bvh::build(std::move(small), 0.f, 1.f)))));
    }
    w.push_back(translate(Vec3(1.5f, 0.4f, 3.5f),
                          std::make_unique<ConstantMedium>(sphere(0.4f, Material::dielectric(1.5f)), 3.0f,
                                                           Material::isotropic(tex_constant(Vec3(0.9f, 0.3f, 0.2f))),
                                                           sc.n_media++)));
}

// Not in the reference's own scenes, but in its type system: ConstantMedium<O> takes any Object as boundary
// (object.rs:533-541).  Shirley's Cornell smoke (two rotated rect_prisms full of smoke, wrappers INSIDE the medium)
// plus a cloud whose boundary is a Bvh of overlapping spheres and prisms, wrapped from outside.
inline void cornell_smoke(Scene& sc, SmallRng& rng) {
    sc.world.list = cornell_box();
    auto& w = sc.world.list;
    Material skin = diffuse_color(splat(0.73f));  // the boundary's own material is never shaded
    w.push_back(std::make_unique<ConstantMedium>(
        translate(Vec3(130.f, 0.f, 65.f), rotate_y(-18.f, rect_prism(Vec3(0.f, 0.f, 0.f), Vec3(165.f, 165.f, 165.f), skin))), 0.01f,
        Material::isotropic(tex_constant(splat(1.f))), sc.n_media++));
    w.push_back(std::make_unique<ConstantMedium>(
        translate(Vec3(265.f, 0.f, 295.f), rotate_y(15.f, rect_prism(Vec3(0.f, 0.f, 0.f), Vec3(165.f, 330.f, 165.f), skin))), 0.01f,
        Material::isotropic(tex_constant(splat(0.f))), sc.n_media++));
    std::vector<ObjectBox> cloud;
    for (int i = 0; i < 6; ++i) cloud.push_back(translate(90.f * rng.gen_vec3(), sphere(45.f, skin)));
    for (int i = 0; i < 3; ++i) {
        Vec3 c = 90.f * rng.gen_vec3();
        cloud.push_back(rect_prism(c, c + Vec3(50.f, 30.f, 40.f), skin));
    }
    w.push_back(translate(Vec3(200.f, 360.f, 200.f),
                          std::make_unique<ConstantMedium>(Bvh::build(std::move(cloud), 0.f, 1.f), 0.02f,
                                                           Material::isotropic(tex_constant(Vec3(0.9f, 0.5f, 0.2f))), sc.n_media++)));
}

// Names are shared with the product's host library (rtiow-rust_b200/csrc/host/scenes.cpp), which
// builds the same scenes from its own code.  top_level_bvh mirrors USE_BVH (main.rs:321,340-351).
inline std::unique_ptr<Scene> build(const std::string& name, uint32_t nx, uint32_t ny, uint64_t scene_seed,
                                    bool top_level_bvh) {
    auto sc = std::make_unique<Scene>();
    sc->perlin = make_perlin_tables(scene_seed);
    SmallRng rng = SmallRng::seed_from_u64(scene_seed);  // main.rs:333
    if (name == "book1") {
        random_scene(*sc, rng, false);
        sc->camera = book1_camera(nx, ny);
        sc->background = Background::SkyGradient;
    } else if (name == "book1_head") {
        random_scene(*sc, rng, true);
        sc->camera = book1_camera(nx, ny);
    } else if (name == "cornell") {
        sc->world.list = cornell_box_with_boxes();
        sc->camera = cornell_camera(nx, ny);
    } else if (name == "cornell_empty") {
        sc->world.list = cornell_box();
        sc->camera = cornell_camera(nx, ny);
    } else if (name == "bench_cornell") {  // benches/scene.rs: Cornell+boxes seen by the book-1 camera
        sc->world.list = cornell_box_with_boxes();
        sc->camera = book1_camera(nx, ny);
    } else if (name == "final") {
        book_final_scene(*sc, rng);
        sc->camera = Camera::look(Vec3(478.f, 278.f, -600.f), Vec3(278.f, 278.f, 0.f), Vec3(0.f, 1.f, 0.f), 40.f,
                                  static_cast<float>(nx) / static_cast<float>(ny), 0.0f, 10.f, 0.f, 1.f);
    } else if (name == "motion_test") {
        motion_test(*sc);
        sc->camera = cornell_camera(nx, ny);
    } else if (name == "volume_test") {
        volume_test(*sc);
        sc->camera = cornell_camera(nx, ny);
    } else if (name == "simple_light") {
        simple_light_scene(*sc, rng);
        sc->camera = cornell_camera(nx, ny);
    } else if (name == "cornell_smoke") {
        cornell_smoke(*sc, rng);
        sc->camera = cornell_camera(nx, ny);
    } else if (name == "kitchen_sink") {
        kitchen_sink(*sc, rng);
        sc->camera = Camera::look(Vec3(6.f, 3.f, 8.f), Vec3(0.f, 1.f, 0.f), Vec3(0.f, 1.f, 0.f), 35.f,
                                  static_cast<float>(nx) / static_cast<float>(ny), 0.05f, 10.f, 0.f, 1.f);
    } else {
        throw std::runtime_error("unknown scene: " + name);
    }
    if (top_level_bvh) {  // main.rs:340-345
        sc->world.bvh = Bvh::build(std::move(sc->world.list), 0.f, 1.f);
        sc->world.list.clear();
    }
    return sc;
}

}  // namespace scenes
}  // namespace oracle
