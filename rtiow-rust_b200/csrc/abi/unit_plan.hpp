// How a launch's samples are cut into work units and its tiles into strips (KParams in path_logic.cuh).  Host-only,
// shared by the ABI (enqueue_render) and tests/kernel_host_harness.cpp, which walks every unit of a plan on the CPU.
#pragma once
#include <algorithm>
#include <cstdint>

#include "../device/path_logic.cuh"

namespace rtiow {

// Strips of 2^shift consecutive tiles: the smallest shift with at most max_strips strips.
inline void plan_strips(uint32_t n_groups, uint32_t max_strips, uint32_t* shift, uint32_t* n_strips) {
    uint32_t sh = 0;
    while (((n_groups + (1u << sh) - 1u) >> sh) > std::max(1u, max_strips)) ++sh;
    *shift = sh;
    *n_strips = (n_groups + (1u << sh) - 1u) >> sh;
}

// Work units of one pass of s_count samples over n_groups tiles (n_tile_slots = n_strips << shift >= n_groups unit slots
// per chunk: the tiles beyond n_groups are empty).  forced_chunk != 0: one size, that one.  open_scene: sky background.
inline void plan_units(uint32_t n_groups, uint32_t n_tile_slots, uint32_t s_count, uint32_t forced_chunk, bool open_scene,
                       uint64_t resident_warps, KParams& K) {
    K.n_groups = n_groups;
    if (forced_chunk != 0u) {  // forced single size
        K.s_chunk = std::min(forced_chunk, s_count);
        K.s_tail_begin = s_count;
    } else if (open_scene) {
        // open scenes (most units are cheap sky): ONE size — measured on book-1: 10.68 ms against 10.97 ms with a tail
        // of smaller units; closed scenes (Cornell: every path is long) gain 3 % from the tail instead.  Which size: a
        // larger unit saves a share of the whole render (coherent camera rays, fewer atomics), its tail costs a fixed
        // time, so the best size grows with the square root of the work per warp: the largest c of 8, 4, 2, 1 with
        // 4 c^2 <= (one-sample units per resident warp) hits the measured optimum for the full book-1 frame and for a
        // half, a quarter and an eighth of it (profiles/r02/p1_unit_size/)
        const uint64_t units1 = static_cast<uint64_t>(n_groups) * s_count;
        uint32_t c = 8u;
        while (c > 1u && 4ull * c * c * resident_warps > units1) c >>= 1;
        K.s_chunk = std::min(c, s_count);
        K.s_tail_begin = s_count;
    } else {
        // big chunks: the largest of 8, 4, 2, 1 of which every resident warp still gets >= 24 (a unit must stay a
        // small fraction of a warp's share: with 3 units of 8 per warp, 8 GPUs lost 25 % to the unlucky warps)
        const uint32_t body = s_count - (s_count + 4u) / 5u;
        uint32_t big = 8u;
        while (big > 1u && static_cast<uint64_t>(n_groups) * (body / big) < 24ull * resident_warps) big >>= 1;
        K.s_chunk = std::min(big, s_count);
        // the tail: about a fifth of the samples in chunks of at most half that size, >= 8 per warp
        const uint64_t want_tail_units = 8ull * resident_warps;
        uint32_t tail = (s_count + 4u) / 5u;
        uint32_t small = std::max(1u, K.s_chunk / 2u);
        while (small > 1u && static_cast<uint64_t>(n_groups) * (tail / small) < want_tail_units) small >>= 1;
        uint32_t big_samples = (s_count - tail) / K.s_chunk * K.s_chunk;  // whole big chunks
        if (K.s_chunk == 1u) big_samples = s_count;                        // nothing smaller to end with
        K.s_tail_begin = big_samples;
        K.s_chunk_tail = small;
    }
    K.n_chunks = K.s_tail_begin / std::max(1u, K.s_chunk) + (K.s_tail_begin % std::max(1u, K.s_chunk) ? 1u : 0u);
    if (K.s_tail_begin == s_count) {
        K.n_chunks = (s_count + K.s_chunk - 1) / K.s_chunk;
        K.s_chunk_tail = 1u;
        K.n_chunks_tail = 0u;
    } else {
        K.n_chunks_tail = (s_count - K.s_tail_begin + K.s_chunk_tail - 1) / K.s_chunk_tail;
    }
    if (static_cast<uint64_t>(n_tile_slots) * (K.n_chunks + K.n_chunks_tail) >= (1ull << 32)) {  // keep the unit counter in 32 bits
        K.s_chunk = s_count; K.n_chunks = 1; K.s_tail_begin = s_count; K.n_chunks_tail = 0; K.s_chunk_tail = 1;
    }
    K.n_big_units = n_tile_slots * K.n_chunks;  // (the tiles beyond n_groups, padding of the last strip, are empty units)
    K.n_units = K.n_big_units + n_tile_slots * K.n_chunks_tail;
}

}  // namespace rtiow
