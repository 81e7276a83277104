// Scene builders of the reference, rebuilt on the host mirror's object model:
// src/lib.rs:103-193 (Cornell), src/lib.rs:237-319 (book-1 random_scene, commented out at HEAD and
// written against a deleted API — re-expressed in the live API), src/main.rs:10-319 (test scenes
// and the book-2 final scene), benches/scene.rs:13-30 (the Criterion configuration).
#include <cmath>

#include "rtiow.hpp"

namespace rtiow {

using material::Material;
using object::Box;
using object::StaticX;
using object::StaticY;
using object::StaticZ;

namespace {

Material diffuse_color(Vec3 c) { return Material::Lambertian(texture::constant(c)); }
Box sphere(float r, const Material& m) { return std::make_unique<object::Sphere>(r, m); }
Box translate(Vec3 off, Box o) { return std::make_unique<object::Translate>(off, std::move(o)); }
Box flip(Box o) { return std::make_unique<object::FlipNormals>(std::move(o)); }
Box rect(object::StaticAxis a, float a0, float a1, float b0, float b1, float k, const Material& m) {
    return std::make_unique<object::Rect>(a, Range{a0, a1}, Range{b0, b1}, k, m);
}
Box linear_move(Box o, Vec3 motion) { return std::make_unique<object::LinearMove>(std::move(o), motion); }

camera::Camera cornell_camera(size_t nx, size_t ny, Range exposure) {  // main.rs:12-27
    return camera::Camera::look(Vec3(278.f, 278.f, -800.f), Vec3(278.f, 278.f, 0.f), Vec3(0.f, 1.f, 0.f), 40.f,
                                static_cast<float>(nx) / static_cast<float>(ny), 0.0f, 10.f, exposure);
}
camera::Camera book1_camera(size_t nx, size_t ny, Range exposure) {  // benches/scene.rs:16-30
    return camera::Camera::look(Vec3(13.f, 2.f, 3.f), Vec3(0.f, 0.f, 0.f), Vec3(0.f, 1.f, 0.f), 20.f,
                                static_cast<float>(nx) / static_cast<float>(ny), 0.1f, 10.f, exposure);
}

// lib.rs:238-319.  `head`=false gives the canonical book-1 materials that img/demo-scene.jpg shows
// (grey ground, static spheres, brown Lambertian); `head`=true keeps the commented block's own
// (Perlin ground, moving Lambertians, Perlin light).
std::vector<Box> random_scene(SmallRng& rng, bool head) {
    std::vector<Box> world;
    world.push_back(translate(Vec3(0.f, -1000.f, 0.f),
                              sphere(1000.f, head ? Material::Lambertian(texture::perlin(4.f)) : diffuse_color(Vec3::from(0.5f)))));
    for (int a = -11; a < 11; ++a) {
        for (int b = -11; b < 11; ++b) {
            const float cx = static_cast<float>(a) + 0.9f * rng.gen_f32();
            const float cz = static_cast<float>(b) + 0.9f * rng.gen_f32();
            const Vec3 center(cx, 0.2f, cz);
            if ((center - Vec3(4.f, 0.2f, 0.f)).length() > 0.9f) {
                const float choose_mat = rng.gen_f32();
                if (choose_mat < 0.8f) {
                    const Vec3 c0 = rng.gen_vec3();
                    const Vec3 c1 = rng.gen_vec3();
                    Box s = sphere(0.2f, diffuse_color(c0 * c1));
                    if (head) s = linear_move(std::move(s), Vec3(0.f, rng.gen_range(0.f, 0.5f), 0.f));
                    world.push_back(translate(center, std::move(s)));
                } else if (choose_mat < 0.95f) {
                    const Vec3 albedo = 0.5f * (1.f + rng.gen_vec3());
                    const float fuzz = 0.5f * rng.gen_f32();
                    world.push_back(translate(center, sphere(0.2f, Material::Metal(albedo, fuzz))));
                } else {
                    world.push_back(translate(center, sphere(0.2f, Material::Dielectric(1.5f))));
                }
            }
        }
    }
    world.push_back(translate(Vec3(0.f, 1.f, 0.f), sphere(1.0f, Material::Dielectric(1.5f))));
    world.push_back(translate(Vec3(-4.f, 1.f, 0.f), sphere(1.0f, Material::Metal(Vec3(0.7f, 0.6f, 0.5f), 0.f))));
    world.push_back(translate(Vec3(4.f, 1.f, 0.f),
                              sphere(1.0f, head ? Material::DiffuseLight(texture::perlin(10.f), 4.f)
                                                : diffuse_color(Vec3(0.4f, 0.2f, 0.1f)))));
    return world;
}

std::vector<Box> book_final_scene(SmallRng& rng, Range exposure, uint32_t& n_media) {  // main.rs:161-319
    std::vector<Box> world;
    const Material ground = diffuse_color(Vec3(0.48f, 0.83f, 0.53f));
    {
        std::vector<Box> boxes;
        for (int i = 0; i < 20; ++i) {
            for (int j = 0; j < 20; ++j) {
                const float W = 100.f;
                const Vec3 c0(-1000.f + static_cast<float>(i) * W, 0.f, -1000.f + static_cast<float>(j) * W);
                const Vec3 c1 = c0 + Vec3(W, 100.f * (rng.gen_f32() + 0.01f), W);
                boxes.push_back(object::rect_prism(c0, c1, ground));
            }
        }
        world.push_back(bvh::from_scene(std::move(boxes), exposure));
    }
    world.push_back(rect(StaticY, 123.f, 423.f, 147.f, 412.f, 554.f, Material::DiffuseLight(texture::constant(Vec3::from(1.f)), 7.f)));
    world.push_back(translate(Vec3(400.f, 400.f, 200.f),
                              linear_move(sphere(50.f, diffuse_color(Vec3(0.7f, 0.3f, 0.1f))), Vec3(30.f, 0.f, 0.f))));
    const Material glass = Material::Dielectric(1.5f);
    world.push_back(translate(Vec3(260.f, 150.f, 45.f), sphere(50.f, glass)));
    world.push_back(translate(Vec3(0.f, 150.f, 145.f), sphere(50.f, Material::Metal(Vec3(0.8f, 0.8f, 0.9f), 1.f))));
    world.push_back(translate(Vec3(360.f, 150.f, 145.f), sphere(70.f, glass)));
    world.push_back(std::make_unique<object::ConstantMedium>(translate(Vec3(360.f, 150.f, 145.f), sphere(70.f, glass)), 0.2f,
                                                             Material::Isotropic(texture::constant(Vec3(0.2f, 0.4f, 0.9f))),
                                                             n_media++));
    world.push_back(std::make_unique<object::ConstantMedium>(sphere(5000.f, glass), 0.0001f,
                                                             Material::Isotropic(texture::constant(Vec3::from(1.f))), n_media++));
    world.push_back(translate(Vec3(220.f, 280.f, 300.f), sphere(80.f, Material::Lambertian(texture::perlin(0.05f)))));
    {
        const Material white = diffuse_color(Vec3::from(0.73f));
        std::vector<Box> spheres;
        for (int i = 0; i < 1000; ++i) spheres.push_back(translate(165.f * rng.gen_vec3(), sphere(10.f, white)));
        world.push_back(translate(Vec3(-100.f, 270.f, 395.f), object::rotate_y(15.f, bvh::from_scene(std::move(spheres), exposure))));
    }
    return world;
}

// Not in the reference: reaches every implementor and the nestings its own scenes never build.
std::vector<Box> kitchen_sink(SmallRng& rng, Range exposure, uint32_t& n_media) {
    std::vector<Box> w;
    w.push_back(rect(StaticY, -10.f, 10.f, -10.f, 10.f, 0.f,
                     Material::Lambertian(texture::checker(texture::constant(Vec3(0.2f, 0.3f, 0.1f)), texture::constant(Vec3::from(0.9f))))));
    w.push_back(std::make_unique<object::Scale>(
        Vec3(1.5f, 0.75f, 1.0f), translate(Vec3(-2.f, 1.5f, 0.f), sphere(1.f, Material::Metal(Vec3(0.8f, 0.6f, 0.2f), 0.1f)))));
    w.push_back(object::rotate_y(30.f, translate(Vec3(2.f, 0.f, -1.f),
                                                 object::rect_prism(Vec3(0.f, 0.f, 0.f), Vec3(1.f, 2.f, 1.f), Material::Dielectric(1.5f)))));
    w.push_back(std::make_unique<object::And>(
        translate(Vec3(0.f, 1.f, 2.f), sphere(0.7f, Material::Lambertian(texture::perlin(3.f)))),
        translate(Vec3(0.6f, 1.f, 2.3f), sphere(0.5f, diffuse_color(Vec3(0.7f, 0.1f, 0.1f))))));
    w.push_back(flip(sphere(40.f, Material::DiffuseLight(texture::constant(Vec3(0.6f, 0.7f, 0.9f)), 1.f))));
    w.push_back(translate(Vec3(-1.f, 0.5f, 3.f),
                          linear_move(sphere(0.5f, Material::Lambertian(texture::checker(texture::perlin(5.f),
                                                                                         texture::constant(Vec3(0.1f, 0.1f, 0.8f))))),
                                      Vec3(0.f, 0.5f, 0.f))));
    w.push_back(std::make_unique<object::ConstantMedium>(translate(Vec3(3.f, 1.f, 2.f), sphere(1.f, Material::Dielectric(1.5f))), 0.8f,
                                                         Material::Isotropic(texture::constant(Vec3::from(0.9f))), n_media++));
    w.push_back(translate(Vec3(1.f, 3.f, 0.f),
                          flip(rect(StaticY, -0.5f, 0.5f, -0.5f, 0.5f, 0.f, Material::DiffuseLight(texture::perlin(2.f), 8.f)))));
    w.push_back(translate(Vec3(0.f, 0.2f, 0.f),
                          translate(Vec3(-3.f, 0.f, 3.f), sphere(0.2f, Material::Metal(Vec3(0.9f, 0.9f, 0.9f), 0.9f)))));
    {
        std::vector<Box> small;
        for (int i = 0; i < 8; ++i) {
            const Vec3 off = 2.f * rng.gen_vec3();
            const Vec3 col = rng.gen_vec3();
            small.push_back(translate(off, sphere(0.25f, diffuse_color(col))));
        }
        w.push_back(translate(Vec3(-4.f, 0.f, -2.f),
                              object::rotate_y(-40.f, std::make_unique<object::Scale>(Vec3(1.f, 2.f, 1.f),
                                                                                       bvh::from_scene(std::move(small), exposure)))));
    }
    w.push_back(translate(Vec3(1.5f, 0.4f, 3.5f),
                          std::make_unique<object::ConstantMedium>(sphere(0.4f, Material::Dielectric(1.5f)), 3.0f,
                                                                   Material::Isotropic(texture::constant(Vec3(0.9f, 0.3f, 0.2f))),
                                                                   n_media++)));
    return w;
}

}  // namespace

std::vector<Box> cornell_box();

namespace {
// ConstantMedium<O> with boundaries other than one primitive (object.rs:533-541 takes any Object): Shirley's Cornell
// smoke — two rotated rect_prisms, the wrappers inside the medium — and a cloud bounded by a Bvh, wrapped from outside.
std::vector<Box> cornell_smoke(SmallRng& rng, Range exposure, uint32_t& n_media) {
    std::vector<Box> w = cornell_box();
    const Material skin = diffuse_color(Vec3::from(0.73f));
    w.push_back(std::make_unique<object::ConstantMedium>(
        translate(Vec3(130.f, 0.f, 65.f), object::rotate_y(-18.f, object::rect_prism(Vec3(0.f, 0.f, 0.f), Vec3(165.f, 165.f, 165.f), skin))),
        0.01f, Material::Isotropic(texture::constant(Vec3::from(1.f))), n_media++));
    w.push_back(std::make_unique<object::ConstantMedium>(
        translate(Vec3(265.f, 0.f, 295.f), object::rotate_y(15.f, object::rect_prism(Vec3(0.f, 0.f, 0.f), Vec3(165.f, 330.f, 165.f), skin))),
        0.01f, Material::Isotropic(texture::constant(Vec3::from(0.f))), n_media++));
    std::vector<Box> cloud;
    for (int i = 0; i < 6; ++i) cloud.push_back(translate(90.f * rng.gen_vec3(), sphere(45.f, skin)));
    for (int i = 0; i < 3; ++i) {
        const Vec3 c = 90.f * rng.gen_vec3();
        cloud.push_back(object::rect_prism(c, c + Vec3(50.f, 30.f, 40.f), skin));
    }
    w.push_back(translate(Vec3(200.f, 360.f, 200.f),
                          std::make_unique<object::ConstantMedium>(bvh::from_scene(std::move(cloud), exposure), 0.02f,
                                                                   Material::Isotropic(texture::constant(Vec3(0.9f, 0.5f, 0.2f))),
                                                                   n_media++)));
    return w;
}
}  // namespace

std::vector<Box> cornell_box() {  // lib.rs:103-166
    const Material red = diffuse_color(Vec3(0.65f, 0.05f, 0.05f));
    const Material white = diffuse_color(Vec3::from(0.73f));
    const Material green = diffuse_color(Vec3(0.12f, 0.45f, 0.15f));
    const Material light = Material::DiffuseLight(texture::constant(Vec3::from(1.f)), 15.f);
    std::vector<Box> v;
    v.push_back(rect(StaticY, 213.f, 343.f, 227.f, 332.f, 554.f, light));
    v.push_back(rect(StaticY, 0.f, 555.f, 0.f, 555.f, 0.f, white));          // floor
    v.push_back(flip(rect(StaticZ, 0.f, 555.f, 0.f, 555.f, 555.f, white)));  // rear wall
    v.push_back(flip(rect(StaticY, 0.f, 555.f, 0.f, 555.f, 555.f, white)));  // ceiling
    v.push_back(rect(StaticX, 0.f, 555.f, 0.f, 555.f, 0.f, red));            // right wall
    v.push_back(flip(rect(StaticX, 0.f, 555.f, 0.f, 555.f, 555.f, green)));  // left wall
    return v;
}

std::vector<Box> cornell_box_with_boxes() {  // lib.rs:168-193
    std::vector<Box> scene = cornell_box();
    const Material white = diffuse_color(Vec3::from(0.73f));
    scene.push_back(translate(Vec3(130.f, 0.f, 65.f),
                              object::rotate_y(-18.f, object::rect_prism(Vec3(0.f, 0.f, 0.f), Vec3(165.f, 165.f, 165.f), white))));
    scene.push_back(translate(Vec3(265.f, 0.f, 295.f),
                              object::rotate_y(15.f, object::rect_prism(Vec3(0.f, 0.f, 0.f), Vec3(165.f, 330.f, 165.f), white))));
    return scene;
}

BuiltScene build_scene(const std::string& name, size_t nx, size_t ny, uint64_t scene_seed, bool use_bvh) {
    BuiltScene out;
    out.exposure = Range{0.f, 1.f};
    out.perlin = PerlinTables::generate(scene_seed);
    SmallRng rng = SmallRng::seed_from_u64(scene_seed);  // main.rs:333
    std::vector<Box> world;
    uint32_t n_media = 0;
    if (name == "book1" || name == "book1_head") {
        const bool head = name == "book1_head";
        world = random_scene(rng, head);
        out.camera = book1_camera(nx, ny, out.exposure);
        out.background = head ? Background::Black : Background::SkyGradient;
    } else if (name == "cornell") {
        world = cornell_box_with_boxes();
        out.camera = cornell_camera(nx, ny, out.exposure);
    } else if (name == "cornell_empty") {
        world = cornell_box();
        out.camera = cornell_camera(nx, ny, out.exposure);
    } else if (name == "bench_cornell") {
        world = cornell_box_with_boxes();
        out.camera = book1_camera(nx, ny, out.exposure);
    } else if (name == "final") {
        world = book_final_scene(rng, out.exposure, n_media);
        out.camera = camera::Camera::look(Vec3(478.f, 278.f, -600.f), Vec3(278.f, 278.f, 0.f), Vec3(0.f, 1.f, 0.f), 40.f,
                                          static_cast<float>(nx) / static_cast<float>(ny), 0.0f, 10.f, out.exposure);
    } else if (name == "motion_test") {  // main.rs:33-67
        world = cornell_box();
        world.push_back(translate(Vec3(278.f, 278.f, 278.f),
                                  linear_move(sphere(65.f, diffuse_color(Vec3::from(0.73f))), Vec3(0.f, 100.f, 0.f))));
        out.camera = cornell_camera(nx, ny, out.exposure);
    } else if (name == "volume_test") {  // main.rs:70-108
        world = cornell_box();
        world.push_back(translate(Vec3(278.f, 278.f, 278.f),
                                  std::make_unique<object::ConstantMedium>(
                                      sphere(180.f, diffuse_color(Vec3::from(0.73f))), 0.01f,
                                      Material::Isotropic(texture::constant(Vec3(0.2f, 0.2f, 1.0f))), n_media++)));
        out.camera = cornell_camera(nx, ny, out.exposure);
    } else if (name == "simple_light") {  // main.rs:111-159
        world = cornell_box();
        for (int i = 0; i < 1000; ++i) {
            const Vec3 off = 277.f + 257.f * rng.gen_vec3();
            world.push_back(translate(off, sphere(20.f, diffuse_color(Vec3::from(0.3f)))));
        }
        world.push_back(flip(sphere(1000.f, Material::DiffuseLight(texture::constant(Vec3::from(0.1f)), 1.f))));
        out.camera = cornell_camera(nx, ny, out.exposure);
    } else if (name == "cornell_smoke") {
        world = cornell_smoke(rng, out.exposure, n_media);
        out.camera = cornell_camera(nx, ny, out.exposure);
    } else if (name == "kitchen_sink") {
        world = kitchen_sink(rng, out.exposure, n_media);
        out.camera = camera::Camera::look(Vec3(6.f, 3.f, 8.f), Vec3(0.f, 1.f, 0.f), Vec3(0.f, 1.f, 0.f), 35.f,
                                          static_cast<float>(nx) / static_cast<float>(ny), 0.05f, 10.f, out.exposure);
    } else {
        throw std::runtime_error("unknown scene: " + name);
    }
    if (use_bvh) out.world = std::make_unique<World>(bvh::from_scene(std::move(world), out.exposure));  // main.rs:340-345
    else out.world = std::make_unique<World>(std::move(world));                                          // main.rs:346-350
    return out;
}

}  // namespace rtiow
