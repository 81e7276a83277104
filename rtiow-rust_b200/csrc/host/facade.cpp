// extern "C" facade over the C++ host mirror so that Python (ctypes) can build the named scenes,
// obtain the flattened descriptor + camera that the C ABI consumes, and write PPMs.
#include <cstring>

#include "rtiow.hpp"

namespace {
thread_local std::string g_host_err;
struct HostScene {
    rtiow::BuiltScene built;
    rtiow::SceneBuilder builder;
    const rtiow_scene_desc_t* desc = nullptr;
    rtiow_camera_t camera{};
};
}  // namespace

extern "C" {

const char* rtiow_host_last_error(void) { return g_host_err.c_str(); }

// name: book1 | book1_head | cornell | cornell_empty | bench_cornell | final | motion_test |
//       volume_test | simple_light | kitchen_sink.   use_bvh mirrors USE_BVH (src/main.rs:321).
void* rtiow_host_scene_build(const char* name, uint32_t nx, uint32_t ny, uint64_t scene_seed, int use_bvh) {
    try {
        auto hs = new HostScene();
        try {
            hs->built = rtiow::build_scene(name, nx, ny, scene_seed, use_bvh != 0);
            hs->builder.set_perlin(hs->built.perlin);
            hs->builder.set_background(static_cast<uint32_t>(hs->built.background), rtiow::Vec3(1.f, 1.f, 1.f),
                                       rtiow::Vec3(0.5f, 0.7f, 1.0f));
            hs->built.world->flatten(hs->builder);
            hs->desc = &hs->builder.finish();
            hs->camera = hs->built.camera.to_repr_c();
        } catch (...) {
            delete hs;
            throw;
        }
        return hs;
    } catch (const std::exception& e) {
        g_host_err = e.what();
        return nullptr;
    }
}
void rtiow_host_scene_free(void* h) { delete static_cast<HostScene*>(h); }
const rtiow_scene_desc_t* rtiow_host_scene_desc(void* h) { return static_cast<HostScene*>(h)->desc; }
const rtiow_camera_t* rtiow_host_scene_camera(void* h) { return &static_cast<HostScene*>(h)->camera; }
uint32_t rtiow_host_scene_len(void* h) { return static_cast<uint32_t>(static_cast<HostScene*>(h)->built.world->len()); }

// Camera::look (src/camera.rs:18-50)
int rtiow_host_camera_look(const float* from, const float* at, const float* up, float fov, float aspect, float aperture,
                           float focus_dist, float time0, float time1, rtiow_camera_t* out) {
    if (!from || !at || !up || !out) { g_host_err = "null argument"; return 1; }
    *out = rtiow::camera::Camera::look(rtiow::Vec3(from[0], from[1], from[2]), rtiow::Vec3(at[0], at[1], at[2]),
                                       rtiow::Vec3(up[0], up[1], up[2]), fov, aspect, aperture, focus_dist,
                                       rtiow::Range{time0, time1}).to_repr_c();
    return 0;
}

// print_ppm (src/lib.rs:344-361) into a file; rgb = ny*nx*3 linear floats, row 0 = top.
int rtiow_host_print_ppm(const float* rgb, uint32_t nx, uint32_t ny, const char* path) {
    if (!rgb || !path) { g_host_err = "null argument"; return 1; }
    std::FILE* f = std::fopen(path, "w");
    if (!f) { g_host_err = std::string("cannot open ") + path; return 1; }
    rtiow::Image img;
    img.nx = nx;
    img.ny = ny;
    img.rgb.assign(rgb, rgb + static_cast<size_t>(nx) * ny * 3);
    rtiow::print_ppm(img, f);
    std::fclose(f);
    return 0;
}

// par_cast on a named scene, end to end through the C++ host API (what the driver binary does).
int rtiow_host_par_cast(void* h, uint32_t nx, uint32_t ny, uint32_t ns, uint64_t seed, int device, float* out_rgb) {
    try {
        auto hs = static_cast<HostScene*>(h);
        rtiow::CastOptions o;
        o.seed = seed;
        o.background = hs->built.background;
        o.perlin = hs->built.perlin;
        o.device = device;
        rtiow::Image img = rtiow::par_cast(nx, ny, ns, hs->built.camera, *hs->built.world, o);
        std::memcpy(out_rgb, img.rgb.data(), img.rgb.size() * sizeof(float));
        return 0;
    } catch (const std::exception& e) {
        g_host_err = e.what();
        return 1;
    }
}

}  // extern "C"
