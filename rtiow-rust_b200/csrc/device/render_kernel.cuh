// The sm_100a path-tracing megakernel: one launch runs the reference's whole par_cast body
// (src/lib.rs:363-376) for a block of scanlines and a range of samples:
//   Camera::get_ray (camera.rs:52-63) -> color (lib.rs:60-101) -> hit_top (lib.rs:33-55) ->
//   Bvh::hit / Aabb::hit (bvh.rs:85-120, aabb.rs:18-29) -> Object::hit (object.rs) ->
//   Material::scatter / emitted (material.rs:55-128) -> Texture (texture.rs) -> perlin (perlin.rs)
// The per-path arithmetic lives in path_logic.cuh; this file is the execution model:
//  * persistent grid (a multiple of the SM count); each CTA first stages the whole scene blob
//    into shared memory with TMA bulk copies (cp.async.bulk + mbarrier) when it fits;
//  * one warp lane = one pixel-sample in flight.  A warp owns a group of 32 neighbouring pixels
//    and hands their samples to lanes as they fall idle (__ballot_sync + prefix popcount), so
//    lanes stay busy although path lengths differ wildly; (group, sample chunk) units come from
//    a global counter;
//  * the object tree is a stack-less "threaded" stream in the reference's visiting order, so a
//    lane's traversal state is just a cursor and the nearest hit so far;
//  * per-sample radiance goes to a staging buffer with one 16-byte store; a second kernel folds
//    the samples of each pixel left to right exactly like `.sum()` (vec3.rs:195-203).
#pragma once
#include <type_traits>

#include "path_logic.cuh"

namespace rtiow {

// ------------------------------------------------------------------------------------------------
// TMA staging of the scene blob
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void stage_scene_tma(unsigned char* smem_dst, const unsigned char* gsrc, uint32_t bytes,
                                                uint64_t* mbar) {
    const uint32_t bar = smem_u32(mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        const uint32_t kChunk = 16384;  // a multiple of 128 B; blob sections are 128 B aligned
        for (uint32_t off = 0; off < bytes; off += kChunk) {
            uint32_t n = bytes - off < kChunk ? bytes - off : kChunk;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(smem_dst + off)),
                         "l"(gsrc + off), "r"(n), "r"(bar)
                         : "memory");
        }
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar)
            : "memory");
    }
}

template <bool kSmem, bool kFrames, bool kFast, uint32_t kFeat, int kThreads, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks) render_kernel(const __grid_constant__ KParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;

    using Mem = typename std::conditional<kSmem, MemShared, MemGlobal>::type;
    Mem mem;
#ifdef RT_CTA_TIMELINE
    unsigned long long* const tl = P.cta_times + 6u * blockIdx.x;  // [start, staged, first/last warp out of units, first/last warp exit]
    auto now = [] { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
    if (threadIdx.x == 0) tl[0] = now();
#endif
    // the strip order (KParams) goes behind the scene; the barrier inside stage_scene_tma publishes it
    const uint32_t order_s = smem_u32(smem_raw) + (kSmem ? ((P.blob_bytes + 127u) & ~127u) : 0u);
    for (uint32_t i = threadIdx.x; i < P.n_strips; i += kThreads)
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(order_s + 4u * i), "r"(__ldg(P.work_counter + kOrderHeaderWords + i)) : "memory");
    if constexpr (kSmem) {
        stage_scene_tma(smem_raw, P.blob, P.blob_bytes, &mbar);
        mem.base = smem_u32(smem_raw);
    } else {
        __syncthreads();
        mem.base = P.blob;
    }
#ifdef RT_CTA_TIMELINE
    if (threadIdx.x == 0) tl[1] = now();
#endif
    const SceneT<Mem, kFeat> sc = scene_views<kFeat>(mem, P);

    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;

    // ---- warp-uniform job pool: the samples of one 8x4-pixel tile ------------------------------
    uint32_t pool_xy = 0, pool_pix0 = 0, pool_next = 0, pool_end = 0, pool_s0 = 0;  // pool_xy: first packed row << 16 | first column
    bool exhausted = false;

    // ---- per-lane path state ------------------------------------------------------------------
    bool active = false;
    uint32_t fresh_x = 0, fresh_r = 0;
    PathState st;
    st.pix = 0; st.samp = 0; st.bounce = 0;
    st.ro = splat(0.f); st.rd = splat(0.f); st.strength = splat(1.f);
    st.rtime = 0.f;
    st.rng = Rng{P.key0, P.key1, 0u, 0u};

    for (;;) {
        // ============ 1. hand new pixel-samples to idle lanes (ballot + prefix popcount) ========
        const bool need = !active;
        const uint32_t need_mask = __ballot_sync(0xffffffffu, need);
        bool fresh = false;
        // refill in batches: camera-ray generation costs the warp the same whether 3 or 30 lanes need it
        if (!exhausted && (static_cast<uint32_t>(__popc(need_mask)) >= P.refill_thr || need_mask == 0xffffffffu)) {
            uint32_t my_rank = __popc(need_mask & lt_mask);
            uint32_t remaining = __popc(need_mask);
            for (;;) {
                const uint32_t avail = pool_end - pool_next;
                if (need && !fresh && my_rank < avail) {
                    // job = sample-major over the tile: 32 neighbouring pixels of one sample first
                    const uint32_t job = pool_next + my_rank;
                    const uint32_t x = (pool_xy & 0xffffu) + (job & (kTileW - 1u)), r = (pool_xy >> 16) + ((job >> kTileWLog2) & (kTileH - 1u));  // = tile_pixel()
                    if (x < P.nx && r < P.n_rows) {  // tiles on the right / bottom edge are partly outside
                        st.samp = P.s_begin + pool_s0 + (job >> 5);
                        st.pix = pool_pix0 + (job & 31u);  // tile-major staging: a tile's 32 pixels of one sample are 512 contiguous bytes
                        fresh_x = x;
                        fresh_r = r;
                        fresh = true;
                    }
                }
                const uint32_t taken = remaining < avail ? remaining : avail;
                pool_next += taken;
                remaining -= taken;
                my_rank -= taken;  // only meaningful for lanes still waiting
                if (remaining == 0u) break;
                uint32_t u = 0;
                if (lane == 0) u = atomicAdd(P.work_counter, 1u);
                u = __shfl_sync(0xffffffffu, u, 0);
                if (u >= P.n_units) {
                    exhausted = true;
#ifdef RT_CTA_TIMELINE
                    if (lane == 0) { const unsigned long long t = now(); atomicMin(tl + 2, t); atomicMax(tl + 3, t); }
#endif
                    break;
                }
                // unit u = a chunk of samples of the g-th tile handed out (KParams: two chunk sizes, strip order)
                uint32_t g, s_n;
                unit_samples(P, u, g, pool_s0, s_n);
                {
                    uint32_t strip;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(strip) : "r"(order_s + 4u * (g >> P.order_shift)));
                    g = tile_of_rank(P, g, strip);
                    if (g >= P.n_groups) { pool_next = pool_end = 0u; continue; }  // padding of the last strip
                }
                pool_pix0 = g * 32u;
                pool_xy = tile_origin(P, g);
                pool_next = 0u;
                pool_end = 32u * s_n;
            }
        }
        if (fresh) {
            generate_camera_ray(P, st, fresh_x, fresh_r);
            active = true;
        }
        const bool warp_alive = __ballot_sync(0xffffffffu, active) != 0u;
        // Phase barriers exist only in the kernels for scenes with wrapper frames on subtrees (kFrames): those are
        // the ones whose code outgrows the instruction cache (rtiow_b200.cu, phase_sync).
        const uint32_t bar_id = 1u + (threadIdx.x >> 5) / P.phase_group, bar_threads = P.phase_group * 32u;
        if (kFrames && P.phase_sync != 0u) {
            // the warps of a barrier group enter hit_top together (and leave the kernel together): their
            // instruction fetches then hit the same few KB of code instead of the whole kernel
            uint32_t any;
            asm volatile(
                "{\n\t.reg .pred p, q;\n\t"
                "setp.ne.u32 q, %3, 0;\n\t"
                "bar.red.or.pred p, %1, %2, q;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(any)
                : "r"(bar_id), "r"(bar_threads), "r"(warp_alive ? 1u : 0u)
                : "memory");
            if (any == 0u) break;
        } else if (!warp_alive) {
            break;
        }

        if (kFrames) {
            // ============ 2. World::hit_top ========================================================
            float best_t = 0.f;
            uint32_t best = kNoHit;
            if (active) best = hit_top_stream<kFrames, kFast>(sc, st, best_t);
            if (P.phase_sync == 2u) asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(bar_threads) : "memory");
            // ============ 3. emitted + scatter =====================================================
            if (active) {
                const uint32_t segs = st.bounce + 1u;
                V3 result;
                if (shade_and_scatter(sc, P, st, best, best_t, result)) {
                    P.staging[static_cast<size_t>(st.samp - P.s_begin) * P.npix + st.pix] =
                        make_float4(result.x, result.y, result.z, static_cast<float>(segs));
                    active = false;
                }
            }
        } else if (active) {
            // ============ 2. World::hit_top ========================================================
            float best_t;
            const uint32_t best = hit_top_stream<kFrames, kFast>(sc, st, best_t);
            // ============ 3. emitted + scatter =====================================================
            const uint32_t segs = st.bounce + 1u;
            V3 result;
            if (shade_and_scatter(sc, P, st, best, best_t, result)) {
                // one 16-byte store per pixel-sample (st.global.v4.f32)
                P.staging[static_cast<size_t>(st.samp - P.s_begin) * P.npix + st.pix] =
                    make_float4(result.x, result.y, result.z, static_cast<float>(segs));
                active = false;
            }
        }
    }
#ifdef RT_CTA_TIMELINE
    if (lane == 0) { const unsigned long long t = now(); atomicMin(tl + 4, t); atomicMax(tl + 5, t); }
#endif
}

}  // namespace rtiow
