// rtiow_b200 — the driver that stands in for src/main.rs:321-356, with real flags instead of
// compile-time constants: builds a scene, optionally wraps it in a top-level BVH (USE_BVH), calls
// par_cast on the GPU, reports wall time to stderr and prints the PPM to stdout.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "rtiow.hpp"

int main(int argc, char** argv) {
    std::string scene = "final";  // main.rs:338
    size_t nx = 300, ny = 300, ns = 100;  // main.rs:324-326
    uint64_t seed = 0xDEADBEEFull, scene_seed = 0xDEADBEEFull;  // main.rs:333
    bool use_bvh = false;  // main.rs:321
    int device = 0;
    for (int i = 1; i < argc; ++i) {
        auto val = [&](const char* flag) -> const char* {
            if (std::strcmp(argv[i], flag) == 0 && i + 1 < argc) return argv[++i];
            return nullptr;
        };
        if (const char* v = val("--scene")) scene = v;
        else if (const char* v = val("--nx")) nx = std::strtoull(v, nullptr, 0);
        else if (const char* v = val("--ny")) ny = std::strtoull(v, nullptr, 0);
        else if (const char* v = val("--ns")) ns = std::strtoull(v, nullptr, 0);
        else if (const char* v = val("--seed")) seed = std::strtoull(v, nullptr, 0);
        else if (const char* v = val("--scene-seed")) scene_seed = std::strtoull(v, nullptr, 0);
        else if (const char* v = val("--device")) device = std::atoi(v);
        else if (std::strcmp(argv[i], "--bvh") == 0) use_bvh = true;
        else {
            std::fprintf(stderr, "usage: %s [--scene NAME] [--nx N] [--ny N] [--ns N] [--seed S] [--scene-seed S] [--bvh] [--device D]\n", argv[0]);
            return 2;
        }
    }
    std::fprintf(stderr, "Parallel casting %zu x %zu image using %zux oversampling.\n", nx, ny, ns);
    try {
        rtiow::BuiltScene built = rtiow::build_scene(scene, nx, ny, scene_seed, use_bvh);
        std::fprintf(stderr, use_bvh ? "Generating bounding volume hierarchy.\nDone.\n" : "Testing every ray against every object.\n");
        rtiow::CastOptions o;
        o.seed = seed;
        o.background = built.background;
        o.perlin = built.perlin;
        o.device = device;
        const auto start = std::chrono::steady_clock::now();
        rtiow::Image image = rtiow::par_cast(nx, ny, ns, built.camera, *built.world, o);
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
        std::fprintf(stderr, "Took %.6fs wall time\n", secs);
        rtiow::print_ppm(image, stdout);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
