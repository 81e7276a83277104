// The small kernels around the megakernel: the sample fold, the per-sample export used by parity
// tests, and print_ppm's quantiser.  Included by the ABI translation unit only.
#pragma once
#include "path_logic.cuh"

namespace rtiow {

// ------------------------------------------------------------------------------------------------
// Fold kernel: `(0..ns).map(..).sum()` then `col / ns as f32` (lib.rs:365-374, vec3.rs:195-203).
// One thread per pixel walks its samples in order; staging is [sample][tile][pixel of the tile] (KParams), so a warp
// reads 512 contiguous bytes per sample and writes four 96-byte row pieces.
//
// Where the finished pixel goes: to `n` frames (this GPU's and, through NVLink peer pointers, the
// other GPUs': the fold IS the framebuffer exchange, there is no all-gather behind it), at packed row r
// of the row block — the whole image in a multi-GPU peer render, where every rank folds its own tiles of it.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxFoldDst = 16;
struct FoldDst {
    float* p[kMaxFoldDst];
    uint32_t n;
    TileMap map;
};

__global__ void __launch_bounds__(256) fold_kernel(const float4* __restrict__ staging, float4* __restrict__ accum,
                                                   const __grid_constant__ FoldDst dst, uint32_t npix, uint32_t s_count,
                                                   int first_pass, int last_pass, float ns_f,
                                                   unsigned long long* __restrict__ seg_total,
                                                   uint32_t* __restrict__ strip_longest, uint32_t n_strips, uint32_t order_shift) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    float segs = 0.f, longest = 0.f;
    uint32_t x = 0, r = 0;
    if (p < npix && dst.map.locate(p, x, r)) {
        float4 acc = first_pass ? make_float4(0.f, 0.f, 0.f, 0.f) : accum[p];
        for (uint32_t s = 0; s < s_count; ++s) {
            const float4 v = __ldcs(staging + static_cast<size_t>(s) * npix + p);
            acc.x = acc.x + v.x;
            acc.y = acc.y + v.y;
            acc.z = acc.z + v.z;
            segs += v.w;
            longest = fmaxf(longest, v.w);
        }
        if (last_pass) {
            const size_t o = 3u * (static_cast<size_t>(r) * dst.map.nx + x);
            const float cr = acc.x / ns_f, cg = acc.y / ns_f, cb = acc.z / ns_f;
            for (uint32_t k = 0; k < dst.n; ++k) {
                float* out = dst.p[k] + o;
                out[0] = cr;
                out[1] = cg;
                out[2] = cb;
            }
        } else {
            accum[p] = acc;
        }
    }
    // segments are small integers: exact in f32 up to 2^24 per pixel-pass
    unsigned long long w = static_cast<unsigned long long>(segs);
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if ((threadIdx.x & 31u) == 0u && w) atomicAdd(seg_total, w);
    // A warp is one tile (tile-major staging): its longest path goes into the key of its strip (KParams) — zeroed before the
    // first pass — stored in REVERSE strip order so that the stable sort leaves equal keys bottom strip first.
    if (strip_longest != nullptr) {
        uint32_t m = static_cast<uint32_t>(longest);
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((threadIdx.x & 31u) == 0u && p < npix && m != 0u) atomicMax(strip_longest + (n_strips - 1u - ((p >> 5) >> order_shift)), m);
    }
}

// The default strip order and the values of the strip-order sort: strip n-1, n-2, ..., 0 — the bottom of the row block
// first: the kernel ends when the last path ends, and in the reference's scenes the cheap pixels (sky, one segment) are
// at the top, they make the better tail.  (`ascending`: 0, 1, ..., n-1, for A/B timing.)
__global__ void __launch_bounds__(256) iota_kernel(uint32_t* __restrict__ out, uint32_t n, int ascending) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = ascending ? i : n - 1u - i;
}

// ------------------------------------------------------------------------------------------------
// Cross-GPU hand-shake of the peer-store exchange (rtiow_b200_render_peers): ONE barrier per frame.
// Every rank owns a flag array in its peer-visible allocation; flags[q] is written by rank q only.
//   arrive: after everything enqueued before this kernel on the stream (the fold that stored my rows into every rank's
//           frame), publish `epoch` into MY slot of every rank's array (system-scope release);
//   wait:   spin (system-scope acquire) until every slot of MY array has reached `epoch` — the frame is whole — with a
//           time-out so that a rank that never arrives cannot hang the GPU: *timed_out is set instead.
// One barrier is enough because the frames are double-buffered (rtiow_peer_frame): epoch e is assembled in buffer
// e & 1, so a fold of epoch e + 2 overwrites what its peers read as frame e — and a peer publishes epoch e + 1 only after
// what it enqueued before that call, its reads of frame e included.
// ------------------------------------------------------------------------------------------------
struct PeerFlags {
    unsigned int* p[kMaxFoldDst];  // rank q's flag array (peer pointer)
    uint32_t n;
};

__global__ void peer_barrier_kernel(const __grid_constant__ PeerFlags flags, uint32_t my_rank, unsigned int epoch,
                                    unsigned long long timeout_ns, unsigned int* __restrict__ timed_out) {
    const uint32_t q = threadIdx.x;
    __threadfence_system();  // the frame stores of the kernels before this one, cumulatively
    if (q >= flags.n) return;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags.p[q] + my_rank), "r"(epoch) : "memory");
    const unsigned int* mine = flags.p[my_rank] + q;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        unsigned int v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if (static_cast<int>(v - epoch) >= 0) break;
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) {
            atomicExch(timed_out, 1u);
            break;
        }
        __nanosleep(100);
    }
}

// Per-sample export for parity debugging: staging [s][tile][pixel] -> out [row][x][ns_total][4]
__global__ void __launch_bounds__(256) export_samples_kernel(const float4* __restrict__ staging, float4* __restrict__ out,
                                                             const TileMap map, uint32_t npix, uint32_t s_begin, uint32_t s_count,
                                                             uint32_t ns_total) {
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<size_t>(npix) * s_count) return;
    const uint32_t s = static_cast<uint32_t>(idx / npix), p = static_cast<uint32_t>(idx % npix);
    uint32_t x, r;
    if (!map.locate(p, x, r)) return;
    out[(static_cast<size_t>(r) * map.nx + x) * ns_total + s_begin + s] = staging[idx];
}

// print_ppm's quantiser (lib.rs:344-361): sqrt, then ((255.99 * x) as i32).max(0).min(255)
__global__ void __launch_bounds__(256) ppm_quantise_kernel(const float* __restrict__ in, unsigned char* __restrict__ out, size_t n) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int v = __float2int_rz(255.99f * sqrtf(in[i]));  // saturating, NaN -> 0, like Rust `as i32`
    out[i] = static_cast<unsigned char>(min(max(v, 0), 255));
}

}  // namespace rtiow
