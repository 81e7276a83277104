// TEST INFRASTRUCTURE — never linked into the shipped libraries.
//
// Compiles the megakernel's per-path code (rtiow-rust_b200/csrc/device/path_logic.cuh: camera ray,
// threaded-stream hit_top, shading) for the HOST and runs it pixel-sample by pixel-sample over a
// flattened scene descriptor, with the same blob layout the device uses.  Comparing its output
// with the oracle isolates flattener / stream-logic bugs from CUDA-specific ones, and can be done
// in the CPU-only container.  The GPU parity tests (-m gpu) check the real kernel.
#include <cstdint>
#include <string>
#include <vector>

#include "../rtiow-rust_b200/csrc/abi/scene_blob.hpp"
#include "../rtiow-rust_b200/csrc/device/path_logic.cuh"
#include "../rtiow-rust_b200/csrc/abi/unit_plan.hpp"

// accel = 1: the re-indexed (SAH, ordered) traversal with conservative inner box tests that the device
// uses by default; 2: the same tree with the reference's exact box test at every node; 0: the plain
// reference-order stream.  Bit 9 (0x200) of the callers' `accel` argument: keep rect_prisms as their six Rect items
// instead of fusing them into prism records (scene_blob.hpp fuse_prisms).
static bool g_fuse_prisms = true;
template <uint32_t kFeat>
static int harness_render_impl(const rtiow_scene_desc_t* desc, const rtiow_camera_t* cam, uint32_t nx, uint32_t ny,
                               uint32_t ns, uint64_t seed, uint32_t row_begin, uint32_t row_end, float* out_rgb,
                               float* out_samples, int accel, uint32_t* layout_out, uint32_t row_step, uint32_t row_band) {
    using namespace rtiow;
    bool has_frames = false, uses_perlin = false;
    std::string msg;
    if (int rc = validate_desc(desc, &has_frames, &uses_perlin, &msg)) return rc;
    BlobLayout lay{};
    const std::vector<unsigned char> blob = build_blob(desc, uses_perlin, &lay, accel == 1 ? kBlobFast : (accel == 2 ? kBlobExact : kBlobReferenceOrder), g_fuse_prisms);
    if (layout_out) { layout_out[0] = lay.n_items; layout_out[1] = lay.n_nodes; layout_out[2] = lay.n_accel; layout_out[3] = lay.accel_depth; layout_out[5] = lay.n_prisms; layout_out[6] = static_cast<uint32_t>(blob.size()); layout_out[7] = lay.n_ordered; layout_out[8] = lay.n_derived; }
    KParams P{};
    P.blob = blob.data();
    P.blob_bytes = static_cast<uint32_t>(blob.size());
    P.off_nodes = lay.off_nodes; P.off_frames = lay.off_frames; P.off_ops = lay.off_ops; P.off_mats = lay.off_mats; P.off_tex = lay.off_tex;
    P.off_pvecs = lay.off_pvecs; P.off_pperm = lay.off_pperm; P.off_fnodes = lay.off_fnodes;
    const bool fast = accel == 1;
    std::memcpy(P.cam, cam, sizeof(float) * 21);
    if (row_step == 0) row_step = 1;
    if (row_band == 0) row_band = 1;
    P.nx = nx; P.ny = ny; P.row_begin = row_begin; P.row_step = row_step; P.row_band = row_band;
    P.n_rows = 0;  // bands of row_band rows starting at row_begin, +step, ... clipped to row_end
    for (uint32_t b = row_begin; b < row_end; b += row_step) P.n_rows += row_end - b < row_band ? row_end - b : row_band;
    P.s_begin = 0; P.s_count = ns;
    P.npix = P.n_rows * nx;
    P.key0 = static_cast<uint32_t>(seed); P.key1 = static_cast<uint32_t>(seed >> 32);
    P.bg_kind = desc->background_kind;
    std::memcpy(P.bg0, desc->background_c0, 12);
    std::memcpy(P.bg1, desc->background_c1, 12);
    const SceneT<MemPtr, kFeat> sc = scene_views<kFeat>(MemPtr{blob.data()}, P);
    for (uint32_t pix = 0; pix < P.npix; ++pix) {
        float acc[3] = {0.f, 0.f, 0.f};
        for (uint32_t s = 0; s < ns; ++s) {
            PathState st;
            st.pix = pix; st.samp = s;
            st.rng = Rng{P.key0, P.key1, 0u, 0u};
            generate_camera_ray(P, st, pix % nx, pix / nx);
            V3 result;
            uint32_t segs;
            for (;;) {
                float best_t;
                const uint32_t best = has_frames ? (fast ? hit_top_stream<true, true>(sc, st, best_t) : hit_top_stream<true, false>(sc, st, best_t))
                                                 : (fast ? hit_top_stream<false, true>(sc, st, best_t) : hit_top_stream<false, false>(sc, st, best_t));
                segs = st.bounce + 1u;
                if (shade_and_scatter(sc, P, st, best, best_t, result)) break;
            }
            if (out_samples) {
                float* o = out_samples + (static_cast<size_t>(pix) * ns + s) * 4;
                o[0] = result.x; o[1] = result.y; o[2] = result.z; o[3] = static_cast<float>(segs);
            }
            acc[0] = acc[0] + result.x; acc[1] = acc[1] + result.y; acc[2] = acc[2] + result.z;
        }
        if (out_rgb) {
            const float nsf = static_cast<float>(ns);
            out_rgb[3 * pix] = acc[0] / nsf; out_rgb[3 * pix + 1] = acc[1] / nsf; out_rgb[3 * pix + 2] = acc[2] / nsf;
        }
    }
    return 0;
}

// accel: low byte as above; bit 8 asks for a feature-specialised instantiation of the per-path code (spheres-only or
// rect-list, path_logic.cuh kFeatSpheres / kFeatRects), chosen like the library does; layout_out[4] = which (0 = general).
extern "C" __attribute__((visibility("default"))) int harness_render(const rtiow_scene_desc_t* desc, const rtiow_camera_t* cam, uint32_t nx, uint32_t ny,
                              uint32_t ns, uint64_t seed, uint32_t row_begin, uint32_t row_end, float* out_rgb,
                              float* out_samples, int accel, uint32_t* layout_out, uint32_t row_step, uint32_t row_band) {
    const bool specialise = (accel & 0x100) != 0;
    g_fuse_prisms = (accel & 0x200) == 0;
    accel &= 0xff;
    uint32_t profile = 0;
    if (specialise) {
        bool has_frames = false, uses_perlin = false;
        std::string msg;
        if (int rc = rtiow::validate_desc(desc, &has_frames, &uses_perlin, &msg)) return rc;
        rtiow::BlobLayout lay{};
        rtiow::build_blob(desc, uses_perlin, &lay, accel == 1 ? rtiow::kBlobFast : (accel == 2 ? rtiow::kBlobExact : rtiow::kBlobReferenceOrder));
        const uint32_t needs = rtiow::scene_features(desc) | (lay.n_accel ? static_cast<uint32_t>(rtiow::SF_ACCEL) : 0u) |
                               (lay.n_ordered ? static_cast<uint32_t>(rtiow::SF_ORDERED) : 0u);
        if (!has_frames && (needs & ~rtiow::kFeatSpheres) == 0u) profile = 1;
        else if (!has_frames && (needs & ~rtiow::kFeatRects) == 0u) profile = 2;
        else if ((needs & ~rtiow::kFeatLean) == 0u) profile = 3;
    }
    if (layout_out) layout_out[4] = profile;
    if (profile == 1) return harness_render_impl<rtiow::kFeatSpheres>(desc, cam, nx, ny, ns, seed, row_begin, row_end, out_rgb, out_samples, accel, layout_out, row_step, row_band);
    if (profile == 2) return harness_render_impl<rtiow::kFeatRects>(desc, cam, nx, ny, ns, seed, row_begin, row_end, out_rgb, out_samples, accel, layout_out, row_step, row_band);
    if (profile == 3) return harness_render_impl<rtiow::kFeatLean>(desc, cam, nx, ny, ns, seed, row_begin, row_end, out_rgb, out_samples, accel, layout_out, row_step, row_band);
    return harness_render_impl<rtiow::SF_ALL>(desc, cam, nx, ny, ns, seed, row_begin, row_end, out_rgb, out_samples, accel, layout_out, row_step, row_band);
}

// hit_top for caller-supplied rays: n rays as {ox, oy, oz, dx, dy, dz, time}; out[2*i] = winning item of the
// REFERENCE-ORDER stream (0xffffffff = none, comparable across traversal modes through its primitive record),
// out[2*i+1] = bits of t.  Used to pit the three traversals against each other on adversarial rays.
extern "C" __attribute__((visibility("default"))) int harness_trace_rays(const rtiow_scene_desc_t* desc, uint32_t n, const float* rays,
                                                                         uint32_t* out, int accel) {
    using namespace rtiow;
    g_fuse_prisms = (accel & 0x200) == 0;
    accel &= 0xff;
    bool has_frames = false, uses_perlin = false;
    std::string msg;
    if (int rc = validate_desc(desc, &has_frames, &uses_perlin, &msg)) return rc;
    BlobLayout lay{};
    const std::vector<unsigned char> blob = build_blob(desc, uses_perlin, &lay, accel == 1 ? kBlobFast : (accel == 2 ? kBlobExact : kBlobReferenceOrder), g_fuse_prisms);
    KParams P{};
    P.blob = blob.data();
    P.off_nodes = lay.off_nodes; P.off_frames = lay.off_frames; P.off_ops = lay.off_ops; P.off_mats = lay.off_mats; P.off_tex = lay.off_tex;
    P.off_pvecs = lay.off_pvecs; P.off_pperm = lay.off_pperm; P.off_fnodes = lay.off_fnodes;
    const SceneT<MemPtr, SF_ALL> sc = scene_views<SF_ALL>(MemPtr{blob.data()}, P);
    const bool fast = accel == 1;
    for (uint32_t i = 0; i < n; ++i) {
        PathState st;
        st.pix = 0; st.samp = 0; st.bounce = 0;
        st.rng = Rng{1u, 2u, i, 0u};
        st.ro = mk(rays[7 * i], rays[7 * i + 1], rays[7 * i + 2]);
        st.rd = mk(rays[7 * i + 3], rays[7 * i + 4], rays[7 * i + 5]);
        st.rtime = rays[7 * i + 6];
        st.strength = splat(1.f);
        float best_t = 0.f;
        const uint32_t best = has_frames ? (fast ? hit_top_stream<true, true>(sc, st, best_t) : hit_top_stream<true, false>(sc, st, best_t))
                                         : (fast ? hit_top_stream<false, true>(sc, st, best_t) : hit_top_stream<false, false>(sc, st, best_t));
        // identify the winner by its primitive record (item numbering differs between blobs)
        uint32_t id = 0xffffffffu;
        if (best != kNoHit) {
            const float4 a = sc.item_a(best & kItemMask), b = sc.item_b(best & kItemMask);
            uint32_t h = 2166136261u;
            const bool medium = (f2u(a.w) & 15u) == IT_MEDIUM;  // a[2] of a medium is an item index: differs between blobs
            const uint32_t w[8] = {f2u(a.x), f2u(a.y), medium ? 0u : f2u(a.z), f2u(a.w) & 15u, f2u(b.x), f2u(b.y), f2u(b.z), (f2u(b.w) & ~(FL_BOX_DERIVED << 24)) | (best & ~kItemMask)};  // the library's own flag differs between blobs
            for (uint32_t k = 0; k < 8; ++k) h = (h ^ w[k]) * 16777619u;
            id = h & 0x7fffffffu;
        }
        out[2 * i] = id;
        out[2 * i + 1] = best == kNoHit ? 0u : f2u(best_t);
    }
    return 0;
}

// Debug aid: one pixel-sample's path, segment by segment: out[8*k..] = {ox, oy, oz, dx, dy, dz, bits(t), winner item of THIS blob}.
extern "C" __attribute__((visibility("default"))) int harness_trace_path(const rtiow_scene_desc_t* desc, const rtiow_camera_t* cam, uint32_t nx,
                                                                         uint32_t ny, uint32_t x, uint32_t row, uint32_t sample, uint64_t seed,
                                                                         int accel, float* out, uint32_t max_segments) {
    using namespace rtiow;
    g_fuse_prisms = (accel & 0x200) == 0;
    accel &= 0xff;
    bool has_frames = false, uses_perlin = false;
    std::string msg;
    if (int rc = validate_desc(desc, &has_frames, &uses_perlin, &msg)) return -rc;
    BlobLayout lay{};
    const std::vector<unsigned char> blob = build_blob(desc, uses_perlin, &lay, accel == 1 ? kBlobFast : (accel == 2 ? kBlobExact : kBlobReferenceOrder), g_fuse_prisms);
    KParams P{};
    P.blob = blob.data();
    P.off_nodes = lay.off_nodes; P.off_frames = lay.off_frames; P.off_ops = lay.off_ops; P.off_mats = lay.off_mats; P.off_tex = lay.off_tex;
    P.off_pvecs = lay.off_pvecs; P.off_pperm = lay.off_pperm; P.off_fnodes = lay.off_fnodes;
    std::memcpy(P.cam, cam, sizeof(float) * 21);
    P.nx = nx; P.ny = ny; P.row_begin = 0; P.row_step = 1; P.row_band = 1; P.n_rows = ny; P.npix = nx * ny;
    P.key0 = static_cast<uint32_t>(seed); P.key1 = static_cast<uint32_t>(seed >> 32);
    P.bg_kind = desc->background_kind;
    std::memcpy(P.bg0, desc->background_c0, 12);
    std::memcpy(P.bg1, desc->background_c1, 12);
    const SceneT<MemPtr, SF_ALL> sc = scene_views<SF_ALL>(MemPtr{blob.data()}, P);
    const bool fast = accel == 1;
    PathState st;
    st.pix = row * nx + x; st.samp = sample;
    st.rng = Rng{P.key0, P.key1, 0u, 0u};
    generate_camera_ray(P, st, x, row);
    uint32_t n = 0;
    for (; n < max_segments; ++n) {
        float best_t = 0.f;
        const V3 o = st.ro, d = st.rd;
        const uint32_t best = has_frames ? (fast ? hit_top_stream<true, true>(sc, st, best_t) : hit_top_stream<true, false>(sc, st, best_t))
                                         : (fast ? hit_top_stream<false, true>(sc, st, best_t) : hit_top_stream<false, false>(sc, st, best_t));
        float* r = out + 8 * n;
        r[0] = o.x; r[1] = o.y; r[2] = o.z; r[3] = d.x; r[4] = d.y; r[5] = d.z; r[6] = best == kNoHit ? 0.f : best_t;
        uint32_t id = best;
        if (best != kNoHit) {  // identify the winner by its record
            const float4 a = sc.item_a(best & kItemMask), b = sc.item_b(best & kItemMask);
            uint32_t h = 2166136261u;
            const uint32_t w[6] = {f2u(a.x), f2u(b.x), f2u(b.y), f2u(b.z), f2u(b.w), best >> 28};
            for (uint32_t k = 0; k < 6; ++k) h = (h ^ w[k]) * 16777619u;
            id = h & 0x7fffffffu;
        }
        std::memcpy(&r[7], &id, 4);
        V3 result;
        if (shade_and_scatter(sc, P, st, best, best_t, result)) { ++n; break; }
    }
    return static_cast<int>(n);
}


// Walks every work unit of one launch exactly as the megakernel's refill does (unit_samples, tile_of_rank, tile_origin) with
// the host's own plan (plan_strips, plan_units), and counts how often each (pixel, sample) of the row block is handed out:
// counts[(r * nx + x) * s_count + s].  `order`: the strip order (a permutation of 0 .. n_strips-1), null = bottom first.
// Returns the number of staging slots that the fold's TileMap does NOT map back to the pixel the kernel rendered into
// them, or -1 for bad arguments.  out_plan: {n_groups, n_strips, order_shift, s_chunk, s_chunk_tail, s_tail_begin, n_units}.
extern "C" __attribute__((visibility("default")))
long harness_unit_coverage(uint32_t nx, uint32_t n_rows, uint32_t tile_first, uint32_t tile_step, uint32_t s_count,
                           uint32_t forced_chunk, int open_scene, uint64_t resident_warps, uint32_t max_strips,
                           const uint32_t* order, uint8_t* counts, uint32_t* out_plan) {
    using namespace rtiow;
    if (nx == 0 || n_rows == 0 || tile_step == 0 || s_count == 0 || nx > 65535u || n_rows > 65535u) return -1;
    KParams P{};
    P.nx = nx; P.n_rows = n_rows;
    P.tiles_x = (nx + kTileW - 1u) / kTileW;
    const uint32_t tiles_all = P.tiles_x * ((n_rows + kTileH - 1u) / kTileH);
    if (tile_first >= tiles_all) return 0;
    P.tile_first = tile_first; P.tile_step = tile_step;
    const uint32_t n_groups = (tiles_all - tile_first + tile_step - 1u) / tile_step;
    plan_strips(n_groups, max_strips, &P.order_shift, &P.n_strips);
    P.s_count = s_count;
    plan_units(n_groups, P.n_strips << P.order_shift, s_count, forced_chunk, open_scene != 0, resident_warps, P);
    if (out_plan) {
        const uint32_t v[7] = {n_groups, P.n_strips, P.order_shift, P.s_chunk, P.s_chunk_tail, P.s_tail_begin, P.n_units};
        std::memcpy(out_plan, v, sizeof(v));
    }
    const TileMap map{nx, n_rows, P.tiles_x, tile_first, tile_step};
    long bad = 0;
    for (uint32_t u = 0; u < P.n_units; ++u) {
        uint32_t rank, s0, s_n;
        unit_samples(P, u, rank, s0, s_n);
        const uint32_t strip = order ? order[rank >> P.order_shift] : P.n_strips - 1u - (rank >> P.order_shift);
        const uint32_t g = tile_of_rank(P, rank, strip);
        if (g >= P.n_groups) continue;  // padding of the last strip
        const uint32_t xy = tile_origin(P, g);
        for (uint32_t job = 0; job < 32u * s_n; ++job) {
            const uint32_t x = (xy & 0xffffu) + (job & (kTileW - 1u)), r = (xy >> 16) + ((job >> kTileWLog2) & (kTileH - 1u));
            if (!(x < nx && r < n_rows)) continue;
            uint32_t fx = 0, fr = 0;
            if (!map.locate(g * 32u + (job & 31u), fx, fr) || fx != x || fr != r) ++bad;
            uint8_t& c = counts[(static_cast<size_t>(r) * nx + x) * s_count + s0 + (job >> 5)];
            if (c < 255) ++c;
        }
    }
    // staging slots of edge tiles that lie outside the block must be skipped by the fold as well
    for (uint32_t p = 0; p < n_groups * 32u; ++p) {
        uint32_t fx = 0, fr = 0;
        const uint32_t xy = tile_origin(P, p >> 5);
        const uint32_t x = (xy & 0xffffu) + (p & (kTileW - 1u)), r = (xy >> 16) + ((p >> kTileWLog2) & (kTileH - 1u));
        if (map.locate(p, fx, fr) != (x < nx && r < n_rows)) ++bad;
    }
    return bad;
}
