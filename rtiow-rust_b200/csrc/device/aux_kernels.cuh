// The small kernels around the megakernel: the sample fold, the per-sample export used by parity
// tests, and print_ppm's quantiser.  Included by the ABI translation unit only.
#pragma once
#include "rt_math.cuh"

namespace rtiow {

// ------------------------------------------------------------------------------------------------
// Fold kernel: `(0..ns).map(..).sum()` then `col / ns as f32` (lib.rs:365-374, vec3.rs:195-203).
// One thread per pixel walks its samples in order; staging is [sample][pixel] so a warp reads
// 512 contiguous bytes per sample.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fold_kernel(const float4* __restrict__ staging, float4* __restrict__ accum,
                                                   float* __restrict__ out_rgb, uint32_t npix, uint32_t s_count,
                                                   int first_pass, int last_pass, float ns_f,
                                                   unsigned long long* __restrict__ seg_total) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    float segs = 0.f;
    if (p < npix) {
        float4 acc = first_pass ? make_float4(0.f, 0.f, 0.f, 0.f) : accum[p];
        for (uint32_t s = 0; s < s_count; ++s) {
            const float4 v = __ldcs(staging + static_cast<size_t>(s) * npix + p);
            acc.x = acc.x + v.x;
            acc.y = acc.y + v.y;
            acc.z = acc.z + v.z;
            segs += v.w;
        }
        if (last_pass) {
            out_rgb[3u * p + 0u] = acc.x / ns_f;
            out_rgb[3u * p + 1u] = acc.y / ns_f;
            out_rgb[3u * p + 2u] = acc.z / ns_f;
        } else {
            accum[p] = acc;
        }
    }
    // segments are small integers: exact in f32 up to 2^24 per pixel-pass
    unsigned long long w = static_cast<unsigned long long>(segs);
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if ((threadIdx.x & 31u) == 0u && w) atomicAdd(seg_total, w);
}

// Per-sample export for parity debugging: staging [s][pix] -> out [pix][ns_total][4]
__global__ void __launch_bounds__(256) export_samples_kernel(const float4* __restrict__ staging, float4* __restrict__ out,
                                                             uint32_t npix, uint32_t s_begin, uint32_t s_count,
                                                             uint32_t ns_total) {
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<size_t>(npix) * s_count) return;
    const uint32_t s = static_cast<uint32_t>(idx / npix), p = static_cast<uint32_t>(idx % npix);
    out[static_cast<size_t>(p) * ns_total + s_begin + s] = staging[idx];
}

// print_ppm's quantiser (lib.rs:344-361): sqrt, then ((255.99 * x) as i32).max(0).min(255)
__global__ void __launch_bounds__(256) ppm_quantise_kernel(const float* __restrict__ in, unsigned char* __restrict__ out, size_t n) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int v = __float2int_rz(255.99f * sqrtf(in[i]));  // saturating, NaN -> 0, like Rust `as i32`
    out[i] = static_cast<unsigned char>(min(max(v, 0), 255));
}

}  // namespace rtiow
