// Per-path logic of the render megakernel: what ONE lane does for ONE pixel-sample.
//   generate_camera_ray : lib.rs:366-370 + Camera::get_ray (camera.rs:52-63)
//   hit_top_stream      : World::hit_top (lib.rs:33-55) over the flattened traversal stream
//                         = Bvh::hit / Aabb::hit / And::hit / Object::hit in the reference's order
//   shade_and_scatter   : the body of color()'s loop (lib.rs:73-98): emitted + Material::scatter
// The functions are __host__ __device__ so that tests/kernel_host_harness.cpp can run exactly this
// code on the CPU against the oracle (flattener and stream logic can then be debugged without a
// GPU).  The shipped library only ever runs them inside render_kernel (render_kernel.cuh).
#pragma once
#include "rt_math.cuh"

namespace rtiow {

// work tiles of the megakernel: kTileW x kTileH = 32 pixels (render_kernel.cuh); kTileW = 1 << kTileWLog2
#ifndef RT_TILE_W_LOG2
#define RT_TILE_W_LOG2 3
#endif
constexpr uint32_t kTileWLog2 = RT_TILE_W_LOG2, kTileW = 1u << kTileWLog2, kTileH = 32u / kTileW;
constexpr float kNear = 0.001f;               // lib.rs:35,53
constexpr float kF32Max = 3.402823466e+38f;   // std::f32::MAX
constexpr float kF32Min = -3.402823466e+38f;  // std::f32::MIN
constexpr uint32_t kNoHit = 0xffffffffu;

// item kinds / flags mirror include/rtiow_b200.h
enum : uint32_t { IT_END = 0, IT_BBOX = 1, IT_SPHERE = 2, IT_RECT = 3, IT_MEDIUM = 4, IT_SET_FRAME = 5, IT_PRISM = 6, IT_ACCEL = 7 };
// The winning item of a hit_top call is `item | face << 28` (face = which of a prism's six rects won, else 0).
constexpr uint32_t kItemMask = 0x0fffffffu;
constexpr uint32_t kLinkLeafBit = 0x80000000u, kLinkNone = 0x7fffffffu;  // accel_build.hpp
constexpr int kStackDepth = 32;
constexpr uint32_t kOrderHeaderWords = 32u, kMaxStrips = 4096u;  // KParams::work_counter
enum : uint32_t { FL_HAS_OFFSET = 1u, FL_FLIP = 2u, FL_BOX_DERIVED = 16u };  // bits 2..3: rect axis; 16: scene_blob.hpp kFlagBoxDerived
enum : uint32_t { OP_TRANSLATE = 0, OP_SCALE = 1, OP_ROTATE_Y = 2, OP_LINEAR_MOVE = 3, OP_FLIP = 4 };
enum : uint32_t { MAT_LAMBERTIAN = 0, MAT_METAL = 1, MAT_DIELECTRIC = 2, MAT_DIFFUSE_LIGHT = 3, MAT_ISOTROPIC = 4 };
enum : uint32_t { TEX_CONSTANT = 0, TEX_CHECKER = 1, TEX_PERLIN = 2 };

struct KParams {
    const unsigned char* blob;  // device scene blob; items start at offset 0
    uint32_t blob_bytes;
    uint32_t off_nodes, off_frames, off_ops, off_mats, off_tex, off_pvecs, off_pperm;
    uint32_t off_fnodes;        // conservative-test nodes (accel_build.hpp FastNode); 0 = none
    float cam[21];              // rtiow_camera_t
    // packed row r is image row row_begin + (r / row_band) * row_step + r % row_band (n_rows of them)
    uint32_t nx, ny, row_begin, n_rows, row_step, row_band;
    // The n_rows x nx row block is cut into 8x4-pixel tiles, tiles_x per tile row; this launch renders tiles tile_first,
    // tile_first + tile_step, ... (n_groups of them).  One GPU: all of them (0, 1).  Rank r of a G-GPU peer render: (r, G)
    // over the WHOLE frame — every rank's tiles are spread evenly over the image, so the ranks' work differs by well under
    // a per cent where whole bands of rows differed by 6 % (the big spheres of book-1 span only a few band periods).
    uint32_t tile_first, tile_step;
    uint32_t s_begin, s_count;  // samples [s_begin, s_begin + s_count) of every pixel in this pass
    uint32_t npix, tiles_x;     // npix = 32 * n_groups: the staging buffer is tile-major, [sample][tile of this launch][pixel of the tile]
    // Work unit = a chunk of samples of one tile, in two sizes: units [0, n_big_units) are chunks of s_chunk samples
    // covering samples [0, s_tail_begin) of every tile, the rest chunks of s_chunk_tail samples covering the remainder —
    // large units while there is plenty of work (one atomic and one coherent batch of rays per 8 x 32 samples), small ones
    // for the end, because the kernel ends when the last warp finishes its last unit.
    uint32_t s_chunk, n_chunks, n_units;
    uint32_t s_chunk_tail, n_chunks_tail, n_big_units, s_tail_begin, n_groups;
    uint32_t n_strips, order_shift;  // (n_strips << order_shift) >= n_groups: the units of the tiles beyond n_groups are empty
    uint32_t key0, key1;
    uint32_t bg_kind;
    float bg0[3], bg1[3];
    uint32_t refill_thr;        // idle lanes are handed new pixel-samples once this many of a warp wait (>= 1)
    uint32_t phase_sync;        // barriers per round: 1 = before hit_top, 2 = also before shading (code-fetch locality)
    uint32_t phase_group;       // warps per barrier group (divides the CTA's warp count)
    float4* staging;            // [s_count][npix] {r, g, b, segments}
    // The unit counter, and kOrderHeaderWords words behind it the order in which the launch's tiles are handed out: the
    // tiles are grouped into n_strips strips of 2^order_shift consecutive tiles, and the table lists the strips — bottom
    // rows first, or longest paths first once the previous render of the same shape has told where those are
    // (enqueue_render).  Every CTA copies the table (at most kMaxStrips words) into shared memory behind the scene: a
    // lookup in global memory per work unit cost 1.6 % on book-1.
    unsigned int* work_counter;
    unsigned long long* cta_times;  // diagnostic builds only (make EXTRA=-DRT_CTA_TIMELINE=1): six time stamps per CTA
};

// ------------------------------------------------------------------------------------------------
// Work units -> tiles -> pixels (the megakernel's refill, the fold's TileMap and tests/kernel_host_harness.cpp share this)
// ------------------------------------------------------------------------------------------------
// Unit u = samples [s0, s0 + s_n) of the pass, of the rank-th tile handed out.
RT_HD void unit_samples(const KParams& P, uint32_t u, uint32_t& rank, uint32_t& s0, uint32_t& s_n) {
    if (u < P.n_big_units) {
        rank = u / P.n_chunks;
        s0 = (u - rank * P.n_chunks) * P.s_chunk;
        s_n = P.s_tail_begin - s0 < P.s_chunk ? P.s_tail_begin - s0 : P.s_chunk;
    } else {
        const uint32_t v = u - P.n_big_units;
        rank = v / P.n_chunks_tail;
        s0 = P.s_tail_begin + (v - rank * P.n_chunks_tail) * P.s_chunk_tail;
        s_n = P.s_count - s0 < P.s_chunk_tail ? P.s_count - s0 : P.s_chunk_tail;
    }
}
// The rank-th tile handed out is tile (mask - rank % 2^shift) of strip `strip` = order[rank >> shift]: inside a strip the
// tiles go bottom first.  A result >= n_groups is the padding of the last strip.
RT_HD uint32_t tile_of_rank(const KParams& P, uint32_t rank, uint32_t strip) {
    return (strip << P.order_shift) | (~rank & ((1u << P.order_shift) - 1u));
}
// First packed row << 16 | first column of tile g of the launch (tile g * tile_step + tile_first of the row block).
RT_HD uint32_t tile_origin(const KParams& P, uint32_t g) {
    const uint32_t tile = g * P.tile_step + P.tile_first, ty = tile / P.tiles_x;
    return ((ty * kTileH) << 16) | ((tile - ty * P.tiles_x) * kTileW);
}

// The same tiling seen from the staging buffer (fold, sample export): staging slot p = 32 * g + pixel of the tile.
struct TileMap {
    uint32_t nx, n_rows, tiles_x, tile_first, tile_step;
    // staging slot p -> column x and packed row r; false for the pixels of edge tiles that lie outside the block
    RT_HD bool locate(uint32_t p, uint32_t& x, uint32_t& r) const {
        const uint32_t tile = (p >> 5) * tile_step + tile_first, ty = tile / tiles_x;
        x = (tile - ty * tiles_x) * kTileW + (p & (kTileW - 1u));
        r = ty * kTileH + ((p >> kTileWLog2) & (kTileH - 1u));
        return x < nx && r < n_rows;
    }
};

// ------------------------------------------------------------------------------------------------
// Where the scene blob lives.  The per-path code reads it through one of these accessors (byte
// offsets into the blob), so the same source runs against shared memory (ld.shared with a 32-bit
// address: no generic-address arithmetic in the traversal loop), global memory (ld.global.nc, for
// scenes that do not fit in shared memory) and host memory (tests/kernel_host_harness.cpp).
// ------------------------------------------------------------------------------------------------
struct MemPtr {  // generic pointer: host harness
    const unsigned char* base;
    RT_HD float4 ld4(uint32_t off) const { return *reinterpret_cast<const float4*>(base + off); }
    RT_HD uint2 ld2(uint32_t off) const { return *reinterpret_cast<const uint2*>(base + off); }
    RT_HD uint32_t ld1b(uint32_t off) const { return base[off]; }
};

#ifdef __CUDACC__
struct MemShared {
    uint32_t base;  // shared-window address of the staged blob
    __device__ __forceinline__ float4 ld4(uint32_t off) const {
        float4 v;
        asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(base + off));
        return v;
    }
    __device__ __forceinline__ uint2 ld2(uint32_t off) const {
        uint2 v;
        asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(base + off));
        return v;
    }
    __device__ __forceinline__ uint32_t ld1b(uint32_t off) const {
        uint32_t v;
        asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(base + off));
        return v;
    }
};
struct MemGlobal {
    const unsigned char* base;
    __device__ __forceinline__ float4 ld4(uint32_t off) const { return __ldg(reinterpret_cast<const float4*>(base + off)); }
    __device__ __forceinline__ uint2 ld2(uint32_t off) const { return __ldg(reinterpret_cast<const uint2*>(base + off)); }
    __device__ __forceinline__ uint32_t ld1b(uint32_t off) const { return __ldg(base + off); }
};
#endif

// What a scene can contain.  The megakernel is compiled for a few feature masks (csrc/kernels): a kernel built
// for mask M renders every scene whose features are a subset of M, and everything outside M compiles out of it —
// book-1's random_scene (unwrapped spheres, Lambertian / Metal / Dielectric, one re-indexed Bvh) needs 1.8 k
// SASS instructions instead of 4.2 k and no register spills; the Cornell box (rects, two rotated prisms, a
// light) needs no traversal stack at all.  scene_blob.hpp scene_features() computes a scene's mask.
enum : uint32_t {
    SF_SPHERE = 1u,      // Sphere items
    SF_RECT = 2u,        // Rect items
    SF_WRAP = 4u,        // a primitive or medium under a wrapper chain (Translate / Scale / RotateY / LinearMove / Flip op)
    SF_MEDIUM = 8u,      // ConstantMedium
    SF_TEXTURE = 16u,    // checker / Perlin textures
    SF_LIGHT = 32u,      // DiffuseLight
    SF_ISOTROPIC = 64u,  // Isotropic
    SF_SPECULAR = 128u,  // Metal, Dielectric
    SF_ACCEL = 256u,     // re-indexed Bvh subtrees (node / leaf traversal with a per-lane stack)
    SF_MEDIUM_RUN = 512u,  // a ConstantMedium whose boundary is more than one primitive (a rect_prism, a Bvh)
    SF_ORDERED = 1024u,  // ordered leaves behind a re-indexed subtree (scene_blob.hpp)
    SF_CHECKER = 2048u,  // checker texture (f32::sin in double precision)
    SF_SCALE = 4096u,    // Scale wrapper (six IEEE divisions each way)
    SF_ALL = 8191u
};
constexpr uint32_t kFeatSpheres = SF_SPHERE | SF_SPECULAR | SF_ACCEL | SF_ORDERED;  // book-1
constexpr uint32_t kFeatRects = SF_RECT | SF_WRAP | SF_LIGHT;                       // Cornell box
// Everything except what the reference's own scenes never build.  The general kernel is bound by instruction fetch
// (4-5 k SASS instructions against a 32 KB instruction cache): every rare feature that is compiled in costs the final
// scene several percent although it never executes (measured: profiles/r02, bisect of the round-2 additions).
constexpr uint32_t kFeatLean = SF_ALL & ~(SF_MEDIUM_RUN | SF_ORDERED | SF_CHECKER | SF_SCALE);

// Views into the scene blob.
template <class Mem, uint32_t kFeat = SF_ALL>
struct SceneT {
    static constexpr uint32_t feat = kFeat;
    Mem m;
    uint32_t off_nodes, off_frames, off_ops, off_mats, off_tex, off_pvecs, off_pperm, off_fnodes;
    RT_HD float4 item_a(uint32_t i) const { return m.ld4(32u * i); }
    RT_HD float4 item_b(uint32_t i) const { return m.ld4(32u * i + 16u); }
    RT_HD float4 node_q(uint32_t n, uint32_t q) const { return m.ld4(off_nodes + 64u * n + 16u * q); }
    RT_HD uint2 frame(uint32_t f) const { return m.ld2(off_frames + 8u * f); }
    RT_HD float4 op(uint32_t k) const { return m.ld4(off_ops + 16u * k); }
    RT_HD float4 mat(uint32_t id, uint32_t half) const { return m.ld4(off_mats + 32u * id + 16u * half); }
    RT_HD float4 tex(uint32_t id, uint32_t half) const { return m.ld4(off_tex + 32u * id + 16u * half); }
    RT_HD float4 pvec(uint32_t i) const { return m.ld4(off_pvecs + 16u * i); }
    RT_HD uint32_t pperm(uint32_t i) const { return m.ld1b(off_pperm + i); }
};

template <uint32_t kFeat, class Mem>
RT_HD SceneT<Mem, kFeat> scene_views(Mem m, const KParams& P) {
    SceneT<Mem, kFeat> sc;
    sc.m = m;
    sc.off_nodes = P.off_nodes; sc.off_frames = P.off_frames; sc.off_ops = P.off_ops; sc.off_mats = P.off_mats;
    sc.off_tex = P.off_tex; sc.off_pvecs = P.off_pvecs; sc.off_pperm = P.off_pperm;
    sc.off_fnodes = P.off_fnodes;
    return sc;
}

struct Rng {
    uint32_t k0, k1, pixel, sample;
    RT_HD U4 block(uint32_t bounce, uint32_t purpose, uint32_t index) const {
        return philox4x32_10(k0, k1, pixel, sample, (bounce << 16) | purpose, index);
    }
};

// One pixel-sample in flight.
struct PathState {
    uint32_t pix, samp, bounce;
    V3 ro, rd, strength;
    float rtime;
    Rng rng;
};

// What hit_top needs to know about the path besides the ray it is tracing, read on demand
// (wrapper frames need the ray's time, a ConstantMedium its random numbers): a view, so that a
// caller can keep those fields wherever it likes.
struct PathStateView {
    const PathState* st;
    RT_HD V3 ro() const { return st->ro; }
    RT_HD V3 rd() const { return st->rd; }
    RT_HD float rtime() const { return st->rtime; }
    RT_HD Rng rng() const { return st->rng; }
    RT_HD uint32_t bounce() const { return st->bounce; }
};

template <bool kScale>
RT_HD void apply_op_ray(const float4 op, V3& o, V3& d, float time) {
    const uint32_t kind = f2u(op.x);
    const V3 v = mk(op.y, op.z, op.w);
    if (kind == OP_TRANSLATE) {            // object.rs:275-278
        o = o - v;
    } else if (kScale && kind == OP_SCALE) {         // object.rs:309-313
        o = o / v;
        d = d / v;
    } else if (kind == OP_ROTATE_Y) {      // object.rs:357-361
        o = rot_y(o, -v.x, v.y);
        d = rot_y(d, -v.x, v.y);
    } else if (kind == OP_LINEAR_MOVE) {   // object.rs:505-508
        o = o - time * v;
    }
}

template <bool kScale>
RT_HD void apply_op_hit(const float4 op, V3& p, V3& n) {
    const uint32_t kind = f2u(op.x);
    const V3 v = mk(op.y, op.z, op.w);
    if (kind == OP_TRANSLATE) {            // object.rs:279-282
        p = p + v;
    } else if (kScale && kind == OP_SCALE) {         // object.rs:314-318
        p = p * v;
        n = n / v;
    } else if (kind == OP_ROTATE_Y) {      // object.rs:365-369
        p = rot_y(p, v.x, v.y);
        n = rot_y(n, v.x, v.y);
    } else if (kind == OP_FLIP) {          // object.rs:249-252
        n = -n;
    }                                      // LinearMove: result returned unmodified (object.rs:504-511)
}

// The wrapper chain of a frame, applied out of line: wrapped primitives are rare on the hot scenes
// and the Scale/RotateY arithmetic is bulky (IEEE divisions), so one copy of each loop keeps the
// megakernel's instruction footprint small.  ops [first, first + n) given as a byte offset.
struct Ray6 {
    V3 o, d;
};
template <uint32_t kFeat, class Mem>
RT_HD_NOINLINE Ray6 frame_ops_ray(Mem m, uint32_t first_off, uint32_t n, V3 o, V3 d, float time) {
    for (uint32_t k = 0; k < n; ++k) apply_op_ray<(kFeat & SF_SCALE) != 0u>(m.ld4(first_off + 16u * k), o, d, time);  // outermost first
    return Ray6{o, d};
}
template <uint32_t kFeat, class Mem>
RT_HD_NOINLINE Ray6 frame_ops_hit(Mem m, uint32_t first_off, uint32_t n, V3 p, V3 nrm) {
    for (uint32_t k = n; k > 0u; --k) apply_op_hit<(kFeat & SF_SCALE) != 0u>(m.ld4(first_off + 16u * (k - 1u)), p, nrm);  // innermost first
    return Ray6{p, nrm};
}

// Sphere::hit (object.rs:82-111) on a ray already in the sphere's frame.
RT_HD bool sphere_hit_t(V3 o, V3 d, float radius, float t_lo, float t_hi, float& t_out) {
    const float a = dot(d, d);
    const float b = dot(o, d);
    const float c = dot(o, o) - radius * radius;
    const float disc = b * b - a * c;
    if (disc > 0.f) {
        const float sq = sqrtf(disc);
        // A ray leaving this sphere has c ~ 0, so sqrt(disc) rounds to |b| and one numerator is exactly 0:
        // 0 / a = 0 < t_lo is the verdict of the division as well, which would take its slow path for it.
        const bool zero_is_miss = t_lo > 0.f && a > 0.f;
        float num = -b - sq;
        if (!(num == 0.f && zero_is_miss)) {
            const float t = num / a;
            if (t < t_hi && t >= t_lo) { t_out = t; return true; }
        }
        num = -b + sq;
        if (!(num == 0.f && zero_is_miss)) {
            const float t = num / a;
            if (t < t_hi && t >= t_lo) { t_out = t; return true; }
        }
    }
    return false;
}

// Rect<A>::hit (object.rs:183-218) with the axis resolved: oa/da = the ray along the rect's axis, (o1, d1) and
// (o2, d2) along the other two (alphabetical, object.rs:153-181); a = {k, r0.start, r0.end}, b = {r1.start, r1.end}
RT_HD bool rect_hit_axis(float oa, float da, float o1, float d1, float o2, float d2, float4 ia, float4 ib, float t_lo, float t_hi,
                         float& t_out) {
    const float num = ia.x - oa;
    // A ray leaving a rect starts (after rounding) exactly on its plane more often than not: 0 / d = +-0 < t_lo.
    // Same verdict as the division below, without its slow path; d = 0 or NaN must divide (0/0 = NaN is a "hit"
    // in the reference, object.rs:194-197), and so must callers with t_lo <= 0 (ConstantMedium boundaries).
    if (num == 0.f && t_lo > 0.f && da != 0.f && da == da) return false;
    const float t = num / da;
    if (t < t_lo || t >= t_hi) return false;
    const float x = o1 + t * d1;
    const float y = o2 + t * d2;
    if (x < ia.y || x >= ia.z || y < ib.x || y >= ib.y) return false;
    t_out = t;
    return true;
}
// RT_RECT_OUTLINE (set by the translation unit of the lean general kernel, which is bound by instruction fetch): ONE
// out-of-line copy of the rect test, components picked by selects, instead of three axis variants at each of three call
// sites.  Rects are rare in the scenes that kernel renders; their code was 7 % of it.  Final scene 83.6 -> 82.6 ms / 100 spp.
#ifndef RT_RECT_OUTLINE
#define RT_RECT_OUTLINE 0
#endif
#if RT_RECT_OUTLINE
static RT_HD_NOINLINE bool rect_hit_t(V3 o, V3 d, uint32_t axis, float4 ia, float4 ib, float t_lo, float t_hi, float& t_out) {
    const float oa = axis == 0u ? o.x : (axis == 1u ? o.y : o.z), da = axis == 0u ? d.x : (axis == 1u ? d.y : d.z);
    const float o1 = axis == 0u ? o.y : o.x, d1 = axis == 0u ? d.y : d.x;  // "other two" alphabetical (object.rs:153-181)
    const float o2 = axis == 2u ? o.y : o.z, d2 = axis == 2u ? d.y : d.z;
    return rect_hit_axis(oa, da, o1, d1, o2, d2, ia, ib, t_lo, t_hi, t_out);
}
#else
RT_HD bool rect_hit_t(V3 o, V3 d, uint32_t axis, float4 ia, float4 ib, float t_lo, float t_hi, float& t_out) {
    // lanes of a warp walk the same item list, so `axis` is uniform in practice: a branch, not selects
    if (axis == 0u) return rect_hit_axis(o.x, d.x, o.y, d.y, o.z, d.z, ia, ib, t_lo, t_hi, t_out);
    if (axis == 1u) return rect_hit_axis(o.y, d.y, o.x, d.x, o.z, d.z, ia, ib, t_lo, t_hi, t_out);
    return rect_hit_axis(o.z, d.z, o.x, d.x, o.y, d.y, ia, ib, t_lo, t_hi, t_out);
}
#endif

// rect_prism (object.rs:420-473) as ONE record {p0, p1}: the six Rect::hit calls of its And tree (object.rs:396-410),
// in the tree's visiting order — z = p1.z, y = p1.y, x = p1.x, then the FlipNormals faces z = p0.z, y = p0.y,
// x = p0.x — each with the arithmetic of rect_hit_axis and the range's end shrunk to the last accepted hit
// (`hit1.or(hit0)`: the last accepted face wins).  `face` = its position in that order.
#ifndef RT_PRISM_INLINE
#define RT_PRISM_INLINE 0
#endif
#if RT_PRISM_INLINE
#define RT_PRISM_FN RT_HD
#else
#define RT_PRISM_FN static RT_HD_NOINLINE  // one copy: the general kernel's code must stay inside the instruction cache
#endif
struct PrismHit {
    float t;
    uint32_t face;  // 0xffffffff: miss
};
RT_PRISM_FN PrismHit prism_hit(V3 o, V3 d, float4 ia, float4 ib, float t_lo, float t_hi);
RT_HD bool prism_hit_t(V3 o, V3 d, float4 ia, float4 ib, float t_lo, float t_hi, float& t_out, uint32_t& face) {
    const PrismHit h = prism_hit(o, d, ia, ib, t_lo, t_hi);
    if (h.face == 0xffffffffu) return false;
    t_out = h.t;
    face = h.face;
    return true;
}
RT_PRISM_FN PrismHit prism_hit(V3 o, V3 d, float4 ia, float4 ib, float t_lo, float t_hi) {
    uint32_t face = 0xffffffffu;
    float t;
    if (rect_hit_axis(o.z, d.z, o.x, d.x, o.y, d.y, make_float4(ib.z, ia.x, ib.x, 0.f), make_float4(ia.y, ib.y, 0.f, 0.f), t_lo, t_hi, t)) { t_hi = t; face = 0u; }
    if (rect_hit_axis(o.y, d.y, o.x, d.x, o.z, d.z, make_float4(ib.y, ia.x, ib.x, 0.f), make_float4(ia.z, ib.z, 0.f, 0.f), t_lo, t_hi, t)) { t_hi = t; face = 1u; }
    if (rect_hit_axis(o.x, d.x, o.y, d.y, o.z, d.z, make_float4(ib.x, ia.y, ib.y, 0.f), make_float4(ia.z, ib.z, 0.f, 0.f), t_lo, t_hi, t)) { t_hi = t; face = 2u; }
    if (rect_hit_axis(o.z, d.z, o.x, d.x, o.y, d.y, make_float4(ia.z, ia.x, ib.x, 0.f), make_float4(ia.y, ib.y, 0.f, 0.f), t_lo, t_hi, t)) { t_hi = t; face = 3u; }
    if (rect_hit_axis(o.y, d.y, o.x, d.x, o.z, d.z, make_float4(ia.y, ia.x, ib.x, 0.f), make_float4(ia.z, ib.z, 0.f, 0.f), t_lo, t_hi, t)) { t_hi = t; face = 4u; }
    if (rect_hit_axis(o.x, d.x, o.y, d.y, o.z, d.z, make_float4(ia.x, ia.y, ib.y, 0.f), make_float4(ia.z, ib.z, 0.f, 0.f), t_lo, t_hi, t)) { t_hi = t; face = 5u; }
    return PrismHit{t_hi, face};
}

// Any primitive item against a ray (o, d) that is already in frame `cur_frame` (whose chain has
// `cur_nops` ops, a prefix of the item's own chain): the item's remaining wrappers are applied
// first, exactly like the nested Object::hit calls.
template <class Mem, uint32_t kFeat, class Path>
RT_HD bool prim_hit_t(const SceneT<Mem, kFeat>& sc, float4 ia, float4 ib, V3 o, V3 d, const Path& path, uint32_t cur_frame,
                      uint32_t cur_nops, float t_lo, float t_hi, float& t_out, uint32_t& face) {
    const uint32_t kind = f2u(ia.w) & 15u;
    const uint32_t frame = f2u(ia.w) >> 4;
    const uint32_t flags = f2u(ib.w) >> 24;
    if (kFeat & SF_WRAP) {
        if (frame != cur_frame) {
            const uint2 fr = sc.frame(frame);
            const Ray6 r = frame_ops_ray<kFeat>(sc.m, sc.off_ops + 16u * (fr.x + cur_nops), fr.y - cur_nops, o, d, path.rtime());
            o = r.o;
            d = r.d;
        }
    }
    if ((kFeat & SF_SPHERE) && (!(kFeat & SF_RECT) || kind == IT_SPHERE)) {
        if (flags & FL_HAS_OFFSET) o = o - mk(ib.x, ib.y, ib.z);
        return sphere_hit_t(o, d, ia.x, t_lo, t_hi, t_out);
    }
    if (kFeat & SF_RECT) {
        if (kind == IT_PRISM) return prism_hit_t(o, d, ia, ib, t_lo, t_hi, t_out, face);
        return rect_hit_t(o, d, (flags >> 2) & 3u, ia, ib, t_lo, t_hi, t_out);
    }
    return false;
}

// ------------------------------------------------------------------------------------------------
// Textures (texture.rs) and Perlin noise (perlin.rs)
// ------------------------------------------------------------------------------------------------
template <class Mem, uint32_t kFeat>
RT_HD_NOINLINE float perlin_noise(const SceneT<Mem, kFeat>& sc, V3 p) {  // perlin.rs:49-64 + trilinear_interp :31-47
    const V3 ijk = mk(floorf(p.x), floorf(p.y), floorf(p.z));
    const V3 uvw = p - ijk;
    const int bi = f2i_rz_sat(ijk.x), bj = f2i_rz_sat(ijk.y), bk = f2i_rz_sat(ijk.z);  // `as i32`
    const V3 uvw3 = uvw * uvw * (splat(3.f) - 2.f * uvw);
    const V3 uvw3_inv = splat(1.f) - uvw3;
    float accum = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const uint32_t ix = sc.pperm((static_cast<uint32_t>(bi) + static_cast<uint32_t>(i)) & 255u);
                const uint32_t iy = sc.pperm(256u + ((static_cast<uint32_t>(bj) + static_cast<uint32_t>(j)) & 255u));
                const uint32_t iz = sc.pperm(512u + ((static_cast<uint32_t>(bk) + static_cast<uint32_t>(k)) & 255u));
                const float4 cv = sc.pvec(ix ^ iy ^ iz);
                const V3 ijkf = mk(static_cast<float>(i), static_cast<float>(j), static_cast<float>(k));
                const float weight = dot(mk(cv.x, cv.y, cv.z), uvw - ijkf);
                const V3 m = ijkf * uvw3 + (splat(1.f) - ijkf) * uvw3_inv;
                accum = accum + ((m.x * m.y) * m.z) * weight;
            }
    return accum;
}

template <class Mem, uint32_t kFeat>
RT_HD_NOINLINE float perlin_turb(const SceneT<Mem, kFeat>& sc, V3 p) {  // perlin.rs:66-75 with depth 7 (texture.rs:24)
    float accum = 0.f, weight = 1.f;
    for (int i = 0; i < 7; ++i) {
        accum += weight * perlin_noise(sc, p);
        weight *= 0.5f;
        p = 2.f * p;
    }
    return fabsf(accum);
}

template <class Mem, uint32_t kFeat>
RT_HD_NOINLINE V3 texture_eval(const SceneT<Mem, kFeat>& sc, uint32_t id, V3 p) {
    for (;;) {
        const float4 t0 = sc.tex(id, 0u);
        const uint32_t kind = f2u(t0.x);
        if (kind == TEX_CONSTANT) return mk(t0.y, t0.z, t0.w);            // texture.rs:8-10
        const float4 t1 = sc.tex(id, 1u);
        if (!(kFeat & SF_CHECKER) || kind == TEX_PERLIN) return splat(perlin_turb(sc, t1.x * p));  // texture.rs:23-26
        const V3 q = 10.f * p;                                            // checker, texture.rs:12-21
        const float s = (sin_f32(q.x) * sin_f32(q.y)) * sin_f32(q.z);
        id = s < 0.f ? f2u(t1.z) : f2u(t1.y);
    }
}

// Material's texture at p; constant textures were baked into the material record at scene upload.
template <class Mem, uint32_t kFeat>
RT_HD V3 material_texture(const SceneT<Mem, kFeat>& sc, float4 m0, float4 m1, V3 p) {
    const uint32_t texkind = (f2u(m0.x) >> 8) & 0xffu;
    if (!(kFeat & SF_TEXTURE) || texkind == TEX_CONSTANT) return mk(m1.x, m1.y, m1.z);
    return texture_eval(sc, f2u(m0.y), p);
}

RT_HD V3 in_unit_sphere(const Rng& rng, uint32_t bounce) {  // vec3.rs:19-26
    for (uint32_t k = 0;; ++k) {
        const U4 w = rng.block(bounce, PURPOSE_SCATTER, k);
        const V3 v = 2.f * mk(unit_f32(w.x), unit_f32(w.y), unit_f32(w.z)) - splat(1.f);
        if (dot(v, v) < 1.f) return v;
    }
}

// ------------------------------------------------------------------------------------------------
// lib.rs:366-370 + Camera::get_ray (camera.rs:52-63) for sample st.samp of pixel st.pix
// (pix counts row-major from the top-left of the rendered row block).
// ------------------------------------------------------------------------------------------------
// The third and later draws of gen_range for the shutter time (camera.rs:55): words of the CAMERA
// blocks 1, 2, ... in order.  Out of line: a draw is rejected only when rounding lands on exposure.end.
static RT_HD_NOINLINE float shutter_time_retry(Rng rng, float scale, float toff, float t_end) {
    for (uint32_t k = 0;; ++k) {
        const U4 tw = rng.block(0u, PURPOSE_CAMERA, 1u + (k >> 2));
        const uint32_t sel = k & 3u;
        const uint32_t word = sel == 0 ? tw.x : (sel == 1 ? tw.y : (sel == 2 ? tw.z : tw.w));
        const float time = f32_1_2(word) * scale + toff;
        if (time < t_end) return time;
    }
}

// (x, r) = column and packed row of st.pix.
RT_HD void generate_camera_ray(const KParams& P, PathState& st, uint32_t x, uint32_t r) {
    const uint32_t band = r / P.row_band;
    const uint32_t y = P.ny - 1u - (P.row_begin + band * P.row_step + (r - band * P.row_band));  // (0..ny).rev()  lib.rs:326-330
    st.rng.pixel = y * P.nx + x;
    st.rng.sample = st.samp;
    const U4 cw = st.rng.block(0u, PURPOSE_CAMERA, 0u);
    const float s = (static_cast<float>(x) + unit_f32(cw.x)) / static_cast<float>(P.nx);  // lib.rs:368
    const float t = (static_cast<float>(y) + unit_f32(cw.y)) / static_cast<float>(P.ny);  // lib.rs:369
    V3 disc;
    for (uint32_t j = 0;; ++j) {  // Vec3::in_unit_disc  vec3.rs:32-39; one Philox block serves two attempts
        const U4 lw = st.rng.block(0u, PURPOSE_LENS, j);
        disc = 2.f * mk(unit_f32(lw.x), unit_f32(lw.y), 0.f) - mk(1.f, 1.f, 0.f);
        if (dot(disc, disc) < 1.f) break;
        disc = 2.f * mk(unit_f32(lw.z), unit_f32(lw.w), 0.f) - mk(1.f, 1.f, 0.f);
        if (dot(disc, disc) < 1.f) break;
    }
    const V3 lens = P.cam[18] * disc;
    const V3 cu = mk(P.cam[12], P.cam[13], P.cam[14]), cv = mk(P.cam[15], P.cam[16], P.cam[17]);
    const V3 offset = lens.x * cu + lens.y * cv;
    // rng.gen_range(exposure.start, exposure.end): rand 0.6.5 UniformFloat::sample_single
    const float scale = P.cam[20] - P.cam[19], toff = P.cam[19] - scale;
    float time = f32_1_2(cw.z) * scale + toff;
    if (!(time < P.cam[20])) {
        time = f32_1_2(cw.w) * scale + toff;
        if (!(time < P.cam[20])) time = shutter_time_retry(st.rng, scale, toff, P.cam[20]);  // a rounding accident, twice in a row
    }
    const V3 corigin = mk(P.cam[0], P.cam[1], P.cam[2]);
    const V3 llc = mk(P.cam[3], P.cam[4], P.cam[5]);
    const V3 hor = mk(P.cam[6], P.cam[7], P.cam[8]), ver = mk(P.cam[9], P.cam[10], P.cam[11]);
    st.ro = corigin + offset;
    st.rd = llc + s * hor + t * ver - corigin - offset;
    st.rtime = time;
    st.strength = splat(1.f);
    st.bounce = 0u;
}

// ------------------------------------------------------------------------------------------------
// Aabb::hit (aabb.rs:18-29) with the ray's 1/d hoisted (the same value at every node).  Returns
// `end > start`; `start` is the entry parameter clamped to NEAR.
// ------------------------------------------------------------------------------------------------
RT_HD bool slab_test_from(float4 mn, float4 mx, V3 fo, V3 inv, float t_start, float t_end);
RT_HD bool slab_test(float4 mn, float4 mx, V3 fo, V3 inv, float t_end, float& start) {
    const float ax = (mn.x - fo.x) * inv.x, ay = (mn.y - fo.y) * inv.y, az = (mn.z - fo.z) * inv.z;
    const float bx = (mx.x - fo.x) * inv.x, by = (mx.y - fo.y) * inv.y, bz = (mx.z - fo.z) * inv.z;
    const float n0x = inv.x < 0.f ? bx : ax, n1x = inv.x < 0.f ? ax : bx;
    const float n0y = inv.y < 0.f ? by : ay, n1y = inv.y < 0.f ? ay : by;
    const float n0z = inv.z < 0.f ? bz : az, n1z = inv.z < 0.f ? az : bz;
    start = rt_max(kNear, rt_max(rt_max(n0x, n0y), n0z));
    const float end = rt_min(t_end, rt_min(rt_min(n1x, n1y), n1z));
    return end > start;
}

// The same test with the caller's t_range.start (a Bvh used as a ConstantMedium boundary is asked with
// f32::MIN.. and hit1.t + 0.0001.., object.rs:551-553).
RT_HD bool slab_test_from(float4 mn, float4 mx, V3 fo, V3 inv, float t_start, float t_end) {
    const float ax = (mn.x - fo.x) * inv.x, ay = (mn.y - fo.y) * inv.y, az = (mn.z - fo.z) * inv.z;
    const float bx = (mx.x - fo.x) * inv.x, by = (mx.y - fo.y) * inv.y, bz = (mx.z - fo.z) * inv.z;
    const float n0x = inv.x < 0.f ? bx : ax, n1x = inv.x < 0.f ? ax : bx;
    const float n0y = inv.y < 0.f ? by : ay, n1y = inv.y < 0.f ? ay : by;
    const float n0z = inv.z < 0.f ? bz : az, n1z = inv.z < 0.f ? az : bz;
    const float start = rt_max(t_start, rt_max(rt_max(n0x, n0y), n0z));
    const float end = rt_min(t_end, rt_min(rt_min(n1x, n1y), n1z));
    return end > start;
}

// The smallest float greater than a finite positive x.
RT_HD float next_up_pos(float x) { return u2f(f2u(x) + 1u); }

// ------------------------------------------------------------------------------------------------
// Traversal state of one lane for one hit_top call.  hit_top is written as resumable steps
// (stream step / node step / leaf step) so that the megakernel can interleave the steps of its 32
// lanes any way it likes (render_kernel.cuh); run back to back they are World::hit_top.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kStreamEnd = 0xffffffffu;

struct Trav {
    uint32_t i;        // next stream item to interpret; kStreamEnd once the stream is finished
    uint32_t cur;      // link being visited inside a re-indexed subtree; kLinkNone outside
    int sp;            // stack height
    float best_t;      // t_range.end so far
    uint32_t best;     // winning item so far
    uint32_t grp_end;  // end of the re-indexed subtree's own items (where its "ordered leaves" start), see trav_stream
    V3 fo, fd, inv;    // the ray in the current BBOX frame and its 1/d (aabb.rs:19)
    uint32_t f_id, f_nops;
    // conservative box test of the re-indexed subtrees (kFast traversal only; see trav_fast_setup)
    V3 fc;             // n = fma(plane, inv, fc)
    uint32_t nbx, nby, nbz;  // blob offsets of the ray's near/far-ordered plane quads of node 0
    float ek2, omek;   // slack of a node with coordinate magnitude m: fma(m, ek2, omek); ek2 = +inf: test nothing
};
// The per-lane stack of (link, entry t) lives apart from Trav: a dynamically indexed array sits in
// local memory, and must not drag the scalars above there with it.
struct TravStack {
    uint32_t link[kStackDepth];
    float t[kStackDepth];
};

RT_HD bool trav_in_node(const Trav& tr) { return tr.cur < kLinkNone; }
RT_HD bool trav_in_leaf(const Trav& tr) { return (tr.cur & kLinkLeafBit) != 0u; }
RT_HD bool trav_done(const Trav& tr) { return tr.cur == kLinkNone && tr.i == kStreamEnd; }

RT_HD void trav_begin(V3 ro, V3 rd, Trav& tr) {
    tr.i = 0u;
    tr.cur = kLinkNone;
    tr.sp = 0;
    tr.best_t = kF32Max;
    tr.best = kNoHit;
    tr.grp_end = 0u;
    tr.fo = ro;
    tr.fd = rd;
    tr.inv = mk(1.f / rd.x, 1.f / rd.y, 1.f / rd.z);
    tr.f_id = 0u;
    tr.f_nops = 0u;
    tr.fc = splat(0.f); tr.nbx = tr.nby = tr.nbz = 0u; tr.ek2 = 0.f; tr.omek = 0.f;
}

// Pop the next link worth visiting: entries that start beyond the current best are dropped (their box test
// `end > start` would fail now, aabb.rs:27-28).  An entry that starts exactly AT the best is still visited: it may
// hold an item that ties with the best and comes earlier in the reference's order (such an item wins there).
RT_HD void trav_pop(Trav& tr, const TravStack& stk) {
    tr.cur = kLinkNone;
    while (tr.sp > 0) {
        --tr.sp;
        if (tr.best_t >= stk.t[tr.sp]) {
            tr.cur = stk.link[tr.sp];
            break;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// A re-indexed Bvh subtree: SAH tree over the reference's leaf boxes, two child
// boxes per node, nearer child first, per-lane stack of (link, entry t).  A leaf is a run of
// primitive items tested in stream order with the reference's arithmetic.  The winner is the
// smallest t, and among equal t the item that comes first in the reference's visiting order, which
// is what `t < t_range.end` with a shrinking end gives in Bvh::hit (bvh.rs:94-106).
// ------------------------------------------------------------------------------------------------
template <class Mem, uint32_t kFeat>
RT_HD void trav_node_step(const SceneT<Mem, kFeat>& sc, Trav& tr, TravStack& stk) {  // requires trav_in_node(tr)
    const uint32_t n = tr.cur;
    const float4 q0 = sc.node_q(n, 0u), q1 = sc.node_q(n, 1u);
    const float4 q2 = sc.node_q(n, 2u), q3 = sc.node_q(n, 3u);
    float s0, s1;
    const float t_end = tr.best == kNoHit ? tr.best_t : next_up_pos(tr.best_t);  // a tie with the best may still win (earlier item)
    const bool h0 = slab_test(q0, q1, tr.fo, tr.inv, t_end, s0);
    const uint32_t l0 = f2u(q0.w), l1 = f2u(q1.w);
    const bool h1 = slab_test(q2, q3, tr.fo, tr.inv, t_end, s1) && l1 != kLinkNone;
    if (h0 && h1) {
        const bool first0 = s0 <= s1;
        stk.link[tr.sp] = first0 ? l1 : l0;
        stk.t[tr.sp] = first0 ? s1 : s0;
        ++tr.sp;
        tr.cur = first0 ? l0 : l1;
    } else if (h0 || h1) {
        tr.cur = h0 ? l0 : l1;
    } else {
        trav_pop(tr, stk);
    }
}

template <class Mem, uint32_t kFeat, class Path>
RT_HD void trav_leaf_test(const SceneT<Mem, kFeat>& sc, const Path& path, Trav& tr, uint32_t link) {
    const uint32_t first = link & 0x00ffffffu, count = (link >> 24) & 0x7fu;
    for (uint32_t j = first; j < first + count; ++j) {
        const float4 ia = sc.item_a(j), ib = sc.item_b(j);
        const float t_hi = (tr.best != kNoHit && j < (tr.best & kItemMask)) ? next_up_pos(tr.best_t) : tr.best_t;
        float t;
        uint32_t face = 0u;
        if (prim_hit_t(sc, ia, ib, tr.fo, tr.fd, path, tr.f_id, tr.f_nops, kNear, t_hi, t, face)) {
            tr.best_t = t;
            tr.best = j | (face << 28);
        }
    }
}
template <class Mem, uint32_t kFeat, class Path>
RT_HD void trav_leaf_step(const SceneT<Mem, kFeat>& sc, const Path& path, Trav& tr, const TravStack& stk) {  // requires trav_in_leaf(tr)
    trav_leaf_test(sc, path, tr, tr.cur);
    trav_pop(tr, stk);
}

// ------------------------------------------------------------------------------------------------
// Conservative box tests for the INNER boxes of a re-indexed subtree.
//
// An inner box only culls: the image cannot change as long as the test never rejects a box whose
// leaf below would pass the reference's Aabb::hit.  So inner boxes (and, as a first filter, leaf
// boxes) are tested with n = fma(plane, 1/d, -(o * 1/d)) on planes that the node stores already
// ordered near/far for the ray's direction signs — 6 FFMA + 4 min/max per box instead of
// 6 FADD + 6 FMUL + 6 FSEL + 4 min/max — and accepted when `end - start > -e2`, where e2 bounds
// twice the difference to the reference's (plane - o) * (1/d):
//     |fl((p - o) * i) - fl(fma(p, i, -fl(o * i)))| <= 4 * 2^-24 * (|p| + |o|) * |i|
// (one rounding of p - o, of the product, of o * i and of the fma, each relative to a quantity
// bounded by (|p| + |o|) * |i|).  e2 = 2^-19 * (m + max|o|) * max|i| is 4x that bound, m being the
// largest |coordinate| in the node.  Leaves then run the reference's exact Aabb::hit on their own
// box before any primitive test (trav_leaf_step_fast), so every accepted hit went through exactly
// the reference's tests.  Axes on which 1/d is not a moderate finite number (d = 0, denormal, inf,
// NaN; |o| huge) switch the ray to "test nothing": e2 = +inf.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kFastNodeBytes = 112u;  // accel_build.hpp FastNode

RT_HD float rt_fma(float a, float b, float c) {
#ifdef __CUDA_ARCH__
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}
RT_HD float rt_max3(float a, float b, float c) { return rt_max(rt_max(a, b), c); }
RT_HD float rt_min3(float a, float b, float c) { return rt_min(rt_min(a, b), c); }

// Call after tr.fo / tr.fd / tr.inv are set (trav_begin, SET_FRAME).
template <class Mem, uint32_t kFeat>
RT_HD void trav_fast_setup(const SceneT<Mem, kFeat>& sc, Trav& tr) {
    const float lo = 7.8886090522101181e-31f, hi = 1.2676506002282294e+30f;  // 2^-100, 2^100
    const float ax = fabsf(tr.inv.x), ay = fabsf(tr.inv.y), az = fabsf(tr.inv.z);
    const float om = rt_max3(fabsf(tr.fo.x), fabsf(tr.fo.y), fabsf(tr.fo.z));
    const bool ok = ax > lo && ax < hi && ay > lo && ay < hi && az > lo && az < hi && om < 67108864.f;  // NaN fails
    tr.nbx = sc.off_fnodes + (tr.inv.x < 0.f ? 16u : 0u);
    tr.nby = sc.off_fnodes + 32u + (tr.inv.y < 0.f ? 16u : 0u);
    tr.nbz = sc.off_fnodes + 64u + (tr.inv.z < 0.f ? 16u : 0u);
    if (ok) {
        tr.fc = mk(-(tr.fo.x * tr.inv.x), -(tr.fo.y * tr.inv.y), -(tr.fo.z * tr.inv.z));
        tr.ek2 = 1.9073486328125e-06f * rt_max3(ax, ay, az);  // 2^-19
        tr.omek = om * tr.ek2;
    } else {  // every fast test passes (0 * plane + 0 = 0 on all axes, slack +inf); exact tests recompute 1/d
        tr.inv = splat(0.f);
        tr.fc = splat(0.f);
        tr.ek2 = u2f(0x7f800000u);
        tr.omek = 0.f;
    }
}

// 1/d for the reference's Aabb::hit (aabb.rs:19): tr.inv, unless the fast setup had to blank it.
template <bool kFast>
RT_HD V3 trav_exact_inv(const Trav& tr) {
    if (kFast && !(tr.ek2 < u2f(0x7f800000u))) return mk(1.f / tr.fd.x, 1.f / tr.fd.y, 1.f / tr.fd.z);
    return tr.inv;
}

template <class Mem, uint32_t kFeat>
RT_HD void trav_node_step_fast(const SceneT<Mem, kFeat>& sc, Trav& tr, TravStack& stk) {  // requires trav_in_node(tr)
    const uint32_t nb = tr.cur * kFastNodeBytes;
    const float4 X = sc.m.ld4(tr.nbx + nb), Y = sc.m.ld4(tr.nby + nb), Z = sc.m.ld4(tr.nbz + nb);  // {near0, far0, near1, far1}
    const float4 L = sc.m.ld4(sc.off_fnodes + 96u + nb);                                            // {link0, link1, m, -}
    const V3 iv = tr.inv;
    const float s0 = rt_max(kNear, rt_max3(rt_fma(X.x, iv.x, tr.fc.x), rt_fma(Y.x, iv.y, tr.fc.y), rt_fma(Z.x, iv.z, tr.fc.z)));
    const float e0 = rt_min(tr.best_t, rt_min3(rt_fma(X.y, iv.x, tr.fc.x), rt_fma(Y.y, iv.y, tr.fc.y), rt_fma(Z.y, iv.z, tr.fc.z)));
    const float s1 = rt_max(kNear, rt_max3(rt_fma(X.z, iv.x, tr.fc.x), rt_fma(Y.z, iv.y, tr.fc.y), rt_fma(Z.z, iv.z, tr.fc.z)));
    const float e1 = rt_min(tr.best_t, rt_min3(rt_fma(X.w, iv.x, tr.fc.x), rt_fma(Y.w, iv.y, tr.fc.y), rt_fma(Z.w, iv.z, tr.fc.z)));
    const float e2 = rt_fma(L.z, tr.ek2, tr.omek);
    const uint32_t l0 = f2u(L.x), l1 = f2u(L.y);
    const bool h0 = (e0 - s0) > -e2;
    const bool h1 = (e1 - s1) > -e2 && l1 != kLinkNone;
    if (h0 && h1) {
        const bool first0 = s0 <= s1;
        stk.link[tr.sp] = first0 ? l1 : l0;
        stk.t[tr.sp] = (first0 ? s1 : s0) - e2;  // popped only while best_t > this: still conservative
        ++tr.sp;
        tr.cur = first0 ? l0 : l1;
    } else if (h0 || h1) {
        tr.cur = h0 ? l0 : l1;
    } else {
        trav_pop(tr, stk);
    }
}

// A leaf of the fast tree: the reference's Aabb::hit (aabb.rs:18-29) on the leaf's own box — the
// BBOX item kept in front of its primitives — and then the primitives.
template <class Mem, uint32_t kFeat, class Path>
RT_HD void trav_leaf_visit_fast(const SceneT<Mem, kFeat>& sc, const Path& path, Trav& tr, uint32_t link) {
    const uint32_t first = link & 0x00ffffffu;
    // The leaf's own box: the BBOX item in front of its primitives, or — a lone Translate{Sphere} / rect_prism, checked at
    // scene upload (scene_blob.hpp leaf_box_is_derivable) — the reference's bounding_box() recomputed from the record.
    const float4 pa = sc.item_a(first), pb = sc.item_b(first);
    const uint32_t pflags = f2u(pb.w) >> 24;
    float4 mn, mx;
    if (pflags & FL_BOX_DERIVED) {
        if ((kFeat & SF_SPHERE) && (!(kFeat & SF_RECT) || (f2u(pa.w) & 15u) == IT_SPHERE)) {
            const bool has = (pflags & FL_HAS_OFFSET) != 0u;  // object.rs:113-118, 285-291
            const float r = pa.x, ox = has ? pb.x : 0.f, oy = has ? pb.y : 0.f, oz = has ? pb.z : 0.f;
            mn = make_float4(has ? -r + ox : -r, has ? -r + oy : -r, has ? -r + oz : -r, 0.f);
            mx = make_float4(has ? r + ox : r, has ? r + oy : r, has ? r + oz : r, 0.f);
        } else {                                              // object.rs:220-233 merged by And (412-416)
            mn = make_float4(pa.x - 0.0001f, pa.y - 0.0001f, pa.z - 0.0001f, 0.f);
            mx = make_float4(pb.x + 0.0001f, pb.y + 0.0001f, pb.z + 0.0001f, 0.f);
        }
    } else {
        mn = sc.item_a(first - 1u);
        mx = sc.item_b(first - 1u);
    }
    float start;
    // t_range.end of the reference's leaf test is the nearest hit of the leaves BEFORE this one (bvh.rs:94-102): a best
    // that comes later in the reference's order may tie with an item of this leaf, which then wins
    const float t_end = (tr.best != kNoHit && first < (tr.best & kItemMask)) ? next_up_pos(tr.best_t) : tr.best_t;
    if (slab_test(mn, mx, tr.fo, trav_exact_inv<true>(tr), t_end, start)) trav_leaf_test(sc, path, tr, link);
}
template <class Mem, uint32_t kFeat, class Path>
RT_HD void trav_leaf_step_fast(const SceneT<Mem, kFeat>& sc, const Path& path, Trav& tr, const TravStack& stk) {  // requires trav_in_leaf(tr)
    trav_leaf_visit_fast(sc, path, tr, tr.cur);
    trav_pop(tr, stk);
}

// Out-of-line copy of prim_hit_t for rare callers (ConstantMedium boundaries).
struct TimeOnlyView;
template <class Mem, uint32_t kFeat>
RT_HD_NOINLINE bool prim_hit_outline(const SceneT<Mem, kFeat> sc, float4 ia, float4 ib, V3 o, V3 d, float time, uint32_t cur_frame,
                                     uint32_t cur_nops, float t_lo, float t_hi, float& t_out);

struct TimeOnlyView {
    float time;
    RT_HD float rtime() const { return time; }
};

template <class Mem, uint32_t kFeat>
RT_HD_NOINLINE bool prim_hit_outline(const SceneT<Mem, kFeat> sc, float4 ia, float4 ib, V3 o, V3 d, float time, uint32_t cur_frame,
                                     uint32_t cur_nops, float t_lo, float t_hi, float& t_out) {
    uint32_t face = 0u;
    return prim_hit_t(sc, ia, ib, o, d, TimeOnlyView{time}, cur_frame, cur_nops, t_lo, t_hi, t_out, face);
}

// ConstantMedium::hit (object.rs:543-575).  Item i is the medium; items [i + 1, run_end) are its boundary object
// flattened like any other object — primitives (one sphere, the rects of a rect_prism ...) and, for a Bvh boundary,
// BBOX items with skip links — in the medium's frame; run_end = bits(a[2]) of the medium item.
//
// boundary.hit(ray, t_lo..t_hi): the run in the reference's visiting order with a shrinking t_range.end
// (Bvh::hit bvh.rs:85-120, And::hit object.rs:396-410).  Only t is needed (object.rs:551-556).
template <class Mem, uint32_t kFeat>
RT_HD_NOINLINE bool boundary_hit(const SceneT<Mem, kFeat> sc, uint32_t first, uint32_t run_end, V3 mo, V3 md, float time,
                                 uint32_t mframe, uint32_t m_nops, float t_lo, float t_hi, float& t_out) {
    bool found = false;
    const V3 inv = mk(1.f / md.x, 1.f / md.y, 1.f / md.z);  // aabb.rs:19
    for (uint32_t j = first; j < run_end;) {
        const float4 ja = sc.item_a(j), jb = sc.item_b(j);
        if ((f2u(ja.w) & 15u) == IT_BBOX) {
            j = slab_test_from(ja, jb, mo, inv, t_lo, t_hi) ? j + 1u : (f2u(ja.w) >> 4);
        } else {
            float t;
            if (prim_hit_outline(sc, ja, jb, mo, md, time, mframe, m_nops, t_lo, t_hi, t)) {
                t_hi = t;
                found = true;
            }
            j += 1u;
        }
    }
    t_out = t_hi;
    return found;
}

struct BestHit {
    float t;
    uint32_t item;
};
template <class Mem, uint32_t kFeat>
RT_HD_NOINLINE BestHit medium_hit(const SceneT<Mem, kFeat> sc, Rng rng, uint32_t bounce, float time, uint32_t i, V3 fo, V3 fd,
                                  uint32_t f_id, uint32_t f_nops, float best_t, uint32_t best) {
    const float4 ia = sc.item_a(i);
    const uint32_t mframe = f2u(ia.w) >> 4;
    V3 mo = fo, md = fd;
    uint32_t m_nops = f_nops;
    if (mframe != f_id) {
        const uint2 fr = sc.frame(mframe);
        const Ray6 r = frame_ops_ray<kFeat>(sc.m, sc.off_ops + 16u * (fr.x + f_nops), fr.y - f_nops, mo, md, time);
        mo = r.o;
        md = r.d;
        m_nops = fr.y;
    }
    const uint32_t run_end = f2u(ia.z);
    float t1, t2;
    bool both;
    const float4 ba = sc.item_a(i + 1u), bb = sc.item_b(i + 1u);
    if (!(kFeat & SF_MEDIUM_RUN) || (run_end == i + 2u && (f2u(ba.w) & 15u) != IT_BBOX)) {  // the usual boundary, one primitive (a sphere of fog): no run to walk
        both = prim_hit_outline(sc, ba, bb, mo, md, time, mframe, m_nops, kF32Min, kF32Max, t1) &&
               prim_hit_outline(sc, ba, bb, mo, md, time, mframe, m_nops, t1 + 0.0001f, kF32Max, t2);
    } else {
        both = boundary_hit(sc, i + 1u, run_end, mo, md, time, mframe, m_nops, kF32Min, kF32Max, t1) &&
               boundary_hit(sc, i + 1u, run_end, mo, md, time, mframe, m_nops, t1 + 0.0001f, kF32Max, t2);
    }
    if (both) {
        t1 = rt_max(t1, kNear);
        t2 = rt_min(t2, best_t);
        if (!(t1 >= t2)) {
            const float len = length(md);
            const float distance_inside = (t2 - t1) * len;
            const U4 mw = rng.block(bounce, PURPOSE_MEDIUM0 + f2u(ia.y), 0u);
            const float hit_distance = -(1.f / ia.x) * ln_f32(unit_f32(mw.x));
            if (hit_distance < distance_inside) {
                best_t = t1 + hit_distance / len;
                best = i;
            }
        }
    }
    return BestHit{best_t, best};
}

// ------------------------------------------------------------------------------------------------
// World::hit_top over the item stream, in the reference's visiting order (re-indexed subtrees are
// order-free inside, see above).
// ------------------------------------------------------------------------------------------------
// Is the best hit so far an item that the reference visits AFTER the ordered leaf being tested (items [q, grp_end) of
// the subtree it was taken out of)?
RT_HD bool ordered_leaf_precedes_best(const Trav& tr, uint32_t q) {
    const uint32_t b = tr.best & kItemMask;
    return tr.best != kNoHit && b >= q && b < tr.grp_end;
}

// Interprets stream items from tr.i until a re-indexed subtree starts (tr.cur = its root) or the
// stream ends (tr.i = kStreamEnd).
template <bool kFrames, bool kFast, class Mem, uint32_t kFeat, class Path>
RT_HD void trav_stream(const SceneT<Mem, kFeat>& sc, const Path& path, Trav& tr) {
    // the last wrapped primitive's frame: the six rects of a rotated prism share one chain
    uint32_t pf_id = tr.f_id;
    V3 po = tr.fo, pd = tr.fd;
    uint32_t i = tr.i;
    uint32_t ord_q = 0u, ord_end = 0u;  // inside an ordered leaf: items [ord_q, tr.grp_end) come after it in the reference's order
    for (;;) {
        const float4 ia = sc.item_a(i);
        const uint32_t kind = f2u(ia.w) & 15u;
        if ((kFeat & SF_ACCEL) && kind == IT_ACCEL) {
            tr.cur = f2u(ia.x);
            tr.sp = 0;
            i = f2u(ia.w) >> 4;
            tr.grp_end = i;
            break;
        } else if (kind == IT_SPHERE || kind == IT_RECT || ((kFeat & SF_RECT) && kind == IT_PRISM)) {
            const float4 ib = sc.item_b(i);
            const uint32_t frame = f2u(ia.w) >> 4;
            if ((kFeat & SF_WRAP) && frame != pf_id) {  // (po, pd) = the ray in this primitive's frame; consecutive items mostly share it
                if (frame == tr.f_id) {
                    po = tr.fo;
                    pd = tr.fd;
                } else {
                    const uint2 fr = sc.frame(frame);
                    const Ray6 r = frame_ops_ray<kFeat>(sc.m, sc.off_ops + 16u * (fr.x + tr.f_nops), fr.y - tr.f_nops, tr.fo, tr.fd, path.rtime());
                    po = r.o;
                    pd = r.d;
                }
                pf_id = frame;
            }
            float t;
            uint32_t face = 0u;
            float t_hi = tr.best_t;
            if ((kFeat & SF_ORDERED) && i < ord_end && ordered_leaf_precedes_best(tr, ord_q)) t_hi = next_up_pos(tr.best_t);
            if (prim_hit_t(sc, ia, ib, po, pd, path, frame, 0u, kNear, t_hi, t, face)) {
                tr.best_t = t;  // nearest = rec.t (lib.rs:42) / t_range.end = h.t (bvh.rs:98-100, object.rs:404-406)
                tr.best = i | (face << 28);
            }
            i += 1u;
        } else if (kind == IT_BBOX) {  // Aabb::hit  aabb.rs:18-29
            const float4 ib = sc.item_b(i);
            float start;
            float t_end = tr.best_t;
            if ((kFeat & SF_ORDERED) && f2u(ib.w) != 0u) {
                // An "ordered leaf" of the re-indexed subtree just walked (scene_blob.hpp): a leaf whose primitive is so
                // large that its computed t can fall outside its own box's computed interval (the radius-1000 ground
                // sphere of book-1), which makes the outcome depend on whether the reference tested it before or after a
                // competing hit.  It is tested here, after the subtree, with the range the reference's order gives it:
                // if the best so far comes LATER in that order, this leaf was asked first there — nothing could cull it,
                // and its hit wins a tie.
                ord_q = f2u(ib.w) - 1u;
                ord_end = f2u(ia.w) >> 4;
                if (ordered_leaf_precedes_best(tr, ord_q)) t_end = kF32Max;
            }
            i = slab_test(ia, ib, tr.fo, trav_exact_inv<kFast>(tr), t_end, start) ? i + 1u : (f2u(ia.w) >> 4);
        } else if ((kFeat & SF_MEDIUM) && kind == IT_MEDIUM) {
            const BestHit h = medium_hit(sc, path.rng(), path.bounce(), path.rtime(), i, tr.fo, tr.fd, tr.f_id, tr.f_nops, tr.best_t, tr.best);
            tr.best_t = h.t;
            tr.best = h.item;
            i = f2u(ia.z);  // past the boundary run
        } else if (kFrames && kind == IT_SET_FRAME) {
            if (kFrames) {
                tr.f_id = f2u(ia.w) >> 4;
                const uint2 fr = sc.frame(tr.f_id);
                const Ray6 r = frame_ops_ray<kFeat>(sc.m, sc.off_ops + 16u * fr.x, fr.y, path.ro(), path.rd(), path.rtime());
                tr.fo = r.o;
                tr.fd = r.d;
                tr.f_nops = fr.y;
                tr.inv = mk(1.f / tr.fd.x, 1.f / tr.fd.y, 1.f / tr.fd.z);
                if (kFast) trav_fast_setup(sc, tr);
                pf_id = tr.f_id;
                po = tr.fo;
                pd = tr.fd;
            }
            i += 1u;
        } else {
            i = kStreamEnd;  // IT_END
            break;
        }
    }
    tr.i = i;
}

// ------------------------------------------------------------------------------------------------
// World::hit_top: the steps above run back to back for one ray.  Returns the index of the winning
// item (kNoHit if none) and its t.
// ------------------------------------------------------------------------------------------------
template <bool kFrames, bool kFast, class Mem, uint32_t kFeat>
RT_HD uint32_t hit_top_stream(const SceneT<Mem, kFeat>& sc, const PathState& st, float& best_t_out) {
    Trav tr;
    TravStack stk;
    const PathStateView path{&st};
    trav_begin(st.ro, st.rd, tr);
    if (kFast) trav_fast_setup(sc, tr);
    for (;;) {
        trav_stream<kFrames, kFast>(sc, path, tr);
        while ((kFeat & SF_ACCEL) && tr.cur != kLinkNone) {  // "while-while": lanes stay together in the cheap node loop
            if (kFast) {
                while (trav_in_node(tr)) trav_node_step_fast(sc, tr, stk);
                if (trav_in_leaf(tr)) trav_leaf_step_fast(sc, path, tr, stk);
            } else {
                while (trav_in_node(tr)) trav_node_step(sc, tr, stk);
                if (trav_in_leaf(tr)) trav_leaf_step(sc, path, tr, stk);
            }
        }
        if (tr.i == kStreamEnd) break;
    }
    best_t_out = tr.best_t;
    return tr.best;
}

// ------------------------------------------------------------------------------------------------
// The body of color()'s loop after hit_top (lib.rs:73-98).  Returns true when the path is finished
// and `result` holds what color() returns; otherwise st carries the scattered ray.
// ------------------------------------------------------------------------------------------------
template <class Mem, uint32_t kFeat>
RT_HD bool shade_and_scatter(const SceneT<Mem, kFeat>& sc, const KParams& P, PathState& st, uint32_t best, float best_t, V3& result) {
    result = splat(0.f);
    if (best == kNoHit) {  // lib.rs:100, or the book-1 sky (rtiow_b200.h RTIOW_BG_SKY_GRADIENT)
        if (P.bg_kind == 1u) {
            const V3 unit_direction = into_unit(st.rd);
            const float t = 0.5f * (unit_direction.y + 1.0f);
            result = st.strength * ((1.0f - t) * mk(P.bg0[0], P.bg0[1], P.bg0[2]) + t * mk(P.bg1[0], P.bg1[1], P.bg1[2]));
        }
        return true;
    }
    // ---- rebuild the HitRecord of the winning item (object.rs:61-71) --------------------------
    const uint32_t face = best >> 28;  // which rect of a prism (0 for everything else)
    best &= kItemMask;
    const float4 ia = sc.item_a(best), ib = sc.item_b(best);
    const uint32_t kind = f2u(ia.w) & 15u;
    const uint32_t flags = f2u(ib.w) >> 24;
    const uint32_t frame = f2u(ia.w) >> 4;
    uint2 fr;
    fr.x = 0u; fr.y = 0u;
    V3 lo = st.ro, ld = st.rd;
    if ((kFeat & SF_WRAP) && frame != 0u) {
        fr = sc.frame(frame);
        const Ray6 r = frame_ops_ray<kFeat>(sc.m, sc.off_ops + 16u * fr.x, fr.y, lo, ld, st.rtime);
        lo = r.o;
        ld = r.d;
    }
    V3 p, n;
    if ((kFeat & SF_SPHERE) && (!(kFeat & (SF_RECT | SF_MEDIUM)) || kind == IT_SPHERE)) {
        if (flags & FL_HAS_OFFSET) lo = lo - mk(ib.x, ib.y, ib.z);
        p = lo + best_t * ld;  // ray.point_at_parameter(t)  object.rs:100
        n = p / ia.x;          // object.rs:104
        if (flags & FL_FLIP) n = -n;
        if (flags & FL_HAS_OFFSET) p = p + mk(ib.x, ib.y, ib.z);
    } else if ((kFeat & SF_RECT) && (!(kFeat & SF_MEDIUM) || kind == IT_RECT || kind == IT_PRISM)) {
        p = lo + best_t * ld;  // object.rs:209
        // a prism's faces in And order: z, y, x at p1, then the three FlipNormals faces at p0 (object.rs:420-473)
        const bool prism = kind == IT_PRISM;
        const uint32_t axis = prism ? 2u - (face >= 3u ? face - 3u : face) : ((flags >> 2) & 3u);
        n = mk(axis == 0 ? 1.f : 0.f, axis == 1 ? 1.f : 0.f, axis == 2 ? 1.f : 0.f);
        if (((flags & FL_FLIP) != 0u) != (prism && face >= 3u)) n = -n;
    } else {                   // medium  object.rs:565-570
        p = lo + best_t * ld;
        n = mk(1.f, 0.f, 0.f);
    }
    if ((kFeat & SF_WRAP) && fr.y != 0u) {
        const Ray6 r = frame_ops_hit<kFeat>(sc.m, sc.off_ops + 16u * fr.x, fr.y, p, n);
        p = r.o;
        n = r.d;
    }

    const uint32_t mat_id = f2u(ib.w) & 0x00ffffffu;
    const float4 m0 = sc.mat(mat_id, 0u), m1 = sc.mat(mat_id, 1u);
    const uint32_t mkind = f2u(m0.x) & 0xffu;
    const V3 rd = st.rd;
    bool done = false;
    if ((kFeat & SF_LIGHT) && mkind == MAT_DIFFUSE_LIGHT) {
        // accum = accum + strength * (brightness * emission(p)); no scatter -> return accum  (lib.rs:76,88-91)
        result = splat(0.f) + st.strength * (m0.z * material_texture(sc, m0, m1, p));
        return true;
    }
    // The scatter draws of this bounce: Vec3::in_unit_sphere (vec3.rs:19-26) for Lambertian, Metal and
    // Isotropic — attempt k reads words x, y, z of SCATTER block k — and word x of block 0 for the
    // Dielectric's reflect-or-refract draw.  One loop, so the Philox rounds exist once in the kernel.
    const bool wants_sphere = !(kFeat & SF_SPECULAR) || mkind != MAT_DIELECTRIC;
    V3 ius = splat(0.f);
    float draw0;
    for (uint32_t k = 0;; ++k) {
        const U4 w = st.rng.block(st.bounce, PURPOSE_SCATTER, k);
        draw0 = unit_f32(w.x);
        if (!wants_sphere) break;
        ius = 2.f * mk(draw0, unit_f32(w.y), unit_f32(w.z)) - splat(1.f);
        if (dot(ius, ius) < 1.f) break;
    }
    if (mkind == MAT_LAMBERTIAN) {             // material.rs:57-65
        const V3 target = p + n + ius;
        st.rd = target - p;
        st.ro = p;
        st.strength = st.strength * material_texture(sc, m0, m1, p);
    } else if ((kFeat & SF_SPECULAR) && mkind == MAT_METAL) {  // material.rs:66-81
        const V3 refl = reflect(into_unit(rd), n);
        st.rd = refl + m0.z * ius;
        st.ro = p;
        if (dot(st.rd, n) > 0.f) st.strength = st.strength * mk(m1.x, m1.y, m1.z);
        else done = true;                      // absorbed: return accum (= 0)
    } else if ((kFeat & SF_SPECULAR) && (!(kFeat & SF_ISOTROPIC) || mkind == MAT_DIELECTRIC)) {  // material.rs:82-107
        const float ref_idx = m0.z;
        V3 outward_normal;
        float ni_over_nt, cosine;
        const float ddn = dot(rd, n);
        if (ddn > 0.f) {
            outward_normal = -n;
            ni_over_nt = ref_idx;
            cosine = ref_idx * ddn / length(rd);
        } else {
            outward_normal = n;
            ni_over_nt = 1.0f / ref_idx;
            cosine = -ddn / length(rd);
        }
        V3 direction;
        bool refracted = refract(rd, outward_normal, ni_over_nt, direction);
        // the draw is consumed only when refract() is Some (material.rs:96-97)
        if (refracted && !(draw0 >= schlick(cosine, ref_idx))) refracted = false;
        if (!refracted) direction = reflect(rd, n);
        st.rd = direction;
        st.ro = p;
        st.strength = st.strength * splat(1.f);
    } else if (kFeat & SF_ISOTROPIC) {         // Isotropic  material.rs:109-116
        st.rd = ius;
        st.ro = p;
        st.strength = st.strength * material_texture(sc, m0, m1, p);
    }
    if (!done) {
        if (st.bounce == 50u) done = true;     // lib.rs:93-95 (accum is 0 here)
        else st.bounce += 1u;
    }
    return done;
}

}  // namespace rtiow
