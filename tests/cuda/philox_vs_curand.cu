// TEST INFRASTRUCTURE — never linked into the shipped libraries.
// The render path's RNG core (rtiow-rust_b200/csrc/device/rt_math.cuh: philox4x32_10) against an independent
// implementation: NVIDIA's curand_Philox4x32_10 (curand_philox4x32_x.h), on n pseudo-random (counter, key)
// pairs.  Prints the number of mismatching words; exit code 0 iff none.
#include <cstdint>
#include <cstdio>
#include <curand_kernel.h>

#include "../../rtiow-rust_b200/csrc/device/rt_math.cuh"

__global__ void compare(uint32_t n, unsigned long long* bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // inputs from a different generator (splitmix-style), so the test does not feed Philox to itself
    unsigned long long z = 0x9E3779B97F4A7C15ull * (i + 1);
    uint32_t w[6];
    for (int k = 0; k < 6; ++k) {
        z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 27; z *= 0x94D049BB133111EBull; z ^= z >> 31;
        w[k] = static_cast<uint32_t>(z >> 16);
    }
    if (i == 0) { for (int k = 0; k < 6; ++k) w[k] = 0u; }
    if (i == 1) { for (int k = 0; k < 6; ++k) w[k] = 0xffffffffu; }
    const rtiow::U4 a = rtiow::philox4x32_10(w[0], w[1], w[2], w[3], w[4], w[5]);
    const uint4 b = curand_Philox4x32_10(make_uint4(w[2], w[3], w[4], w[5]), make_uint2(w[0], w[1]));
    const int d = (a.x != b.x) + (a.y != b.y) + (a.z != b.z) + (a.w != b.w);
    if (d) atomicAdd(bad, static_cast<unsigned long long>(d));
}

int main() {
    const uint32_t n = 1u << 22;
    unsigned long long* d_bad = nullptr;
    unsigned long long bad = ~0ull;
    if (cudaMalloc(&d_bad, 8) != cudaSuccess || cudaMemset(d_bad, 0, 8) != cudaSuccess) { printf("cuda error\n"); return 2; }
    compare<<<(n + 255) / 256, 256>>>(n, d_bad);
    if (cudaMemcpy(&bad, d_bad, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError())); return 2; }
    printf("philox4x32_10 vs curand_Philox4x32_10: %u blocks, %llu mismatching words\n", n, bad);
    return bad == 0 ? 0 : 1;
}
