// Device-side scalar/vector math for the B200 render path of cbiffle/rtiow-rust.
//
// Everything here reproduces the reference's f32 arithmetic operation for operation (same
// association order, no fused multiply-add: the translation unit is compiled with -fmad=false,
// IEEE division and square root), so that a path traced on the device makes exactly the same
// hit/miss and accept/reject decisions as the reference's scalar code.  Citations are to
// /root/reference/src.
#pragma once
#include <cstdint>
#ifndef RT_FAST_MATH
#define RT_FAST_MATH 0  // 1: the tolerance build (Makefile FAST=1); everything below describes the parity build
#endif
#ifdef __CUDACC__
#include <cuda_runtime.h>
#define RT_HD __host__ __device__ __forceinline__
#define RT_HD_NOINLINE __host__ __device__ __noinline__
#else
// Host-only build of the same per-path code, used by tests/kernel_host_harness.cpp to check the
// flattened-stream logic against the oracle without a GPU.  Not part of any shipped library.
#include <cmath>
#include <cstring>
#define RT_HD inline
#define RT_HD_NOINLINE inline
struct float4 { float x, y, z, w; };
struct uint2 { uint32_t x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
#endif

namespace rtiow {

// Bit casts and the few intrinsics the path needs, with host equivalents of identical semantics.
RT_HD uint32_t f2u(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
RT_HD float u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
RT_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32);
#endif
}
RT_HD int f2i_rz_sat(float x) {  // Rust `x as i32`: truncates, saturates, NaN -> 0
#ifdef __CUDA_ARCH__
    return __float2int_rz(x);
#else
    if (x != x) return 0;
    if (x >= 2147483648.0f) return 2147483647;
    if (x <= -2147483648.0f) return -2147483647 - 1;
    return static_cast<int>(x);
#endif
}
RT_HD float rt_max(float a, float b) { return fmaxf(a, b); }  // f32::max: a NaN operand is ignored
RT_HD float rt_min(float a, float b) { return fminf(a, b); }

struct V3 {
    float x, y, z;
};

RT_HD V3 mk(float x, float y, float z) { return V3{x, y, z}; }
RT_HD V3 splat(float v) { return V3{v, v, v}; }                                   // vec3.rs:106-111
RT_HD V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }    // vec3.rs:155-162
RT_HD V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }    // vec3.rs:175-182
RT_HD V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }    // vec3.rs:115-122
RT_HD V3 operator/(V3 a, V3 b) { return V3{a.x / b.x, a.y / b.y, a.z / b.z}; }    // vec3.rs:135-142
RT_HD V3 operator*(float s, V3 v) { return V3{s * v.x, s * v.y, s * v.z}; }       // vec3.rs:125-132
RT_HD V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }       // vec3.rs:145-152
RT_HD V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }                         // vec3.rs:185-192

// zip_with(mul).reduce(add) = (x + y) + z                                                             // vec3.rs:43-46,100-102
RT_HD float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
RT_HD float length(V3 v) { return sqrtf(dot(v, v)); }                             // vec3.rs:59-61
RT_HD V3 into_unit(V3 v) { return v / length(v); }                                // vec3.rs:66-68 (3 divisions)
RT_HD V3 reflect(V3 v, V3 n) { return v - (2.f * dot(v, n)) * n; }                // vec3.rs:313-315

RT_HD bool refract(V3 v, V3 n, float ni_over_nt, V3& out) {                       // vec3.rs:321-330
    V3 uv = into_unit(v);
    float dt = dot(uv, n);
    float discriminant = 1.0f - ni_over_nt * ni_over_nt * (1.f - dt * dt);
    if (discriminant > 0.f) {
        out = ni_over_nt * (uv - dt * n) - sqrtf(discriminant) * n;
        return true;
    }
    return false;
}

// rot() of RotateY::hit                                                                               // object.rs:349-355
RT_HD V3 rot_y(V3 p, float s, float c) {
    return V3{dot(p, mk(c, 0.f, s)), dot(p, mk(0.f, 1.f, 0.f)), dot(p, mk(-s, 0.f, c))};
}

// ---------------------------------------------------------------------------------------------
// Counter-based RNG: Philox4x32-10.  The reference threads `&mut impl Rng` through the path
// (lib.rs:60, camera.rs:52, material.rs:55, object.rs:33); on the device every consumer
// addresses its own word instead: key = seed, counter = (pixel, sample, bounce<<16 | purpose,
// index) — see DESIGN.md "RNG contract".
// ---------------------------------------------------------------------------------------------
struct U4 {
    uint32_t x, y, z, w;
};

RT_HD U4 philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    return U4{c0, c1, c2, c3};
}

enum : uint32_t { PURPOSE_CAMERA = 0, PURPOSE_LENS = 1, PURPOSE_SCATTER = 2, PURPOSE_MEDIUM0 = 16 };

// rng.gen::<f32>(): 24 high bits -> [0,1)   (rand 0.6.5 Standard)
RT_HD float unit_f32(uint32_t w) { return static_cast<float>(w >> 8) * (1.0f / 16777216.0f); }
// rand 0.6.5 UniformFloat: 23 high bits as the mantissa of a float in [1,2)
RT_HD float f32_1_2(uint32_t w) { return u2f(0x3F800000u | (w >> 9)); }

// ---------------------------------------------------------------------------------------------
// Transcendentals of the path, as fixed double-precision algorithms (IEEE +,-,*,/ only, one final
// rounding to f32) so the device and any scalar host implementation of the same contract agree
// bit for bit.  CUDA's logf/sinf/powf would differ from the host libm in the last place, which
// in a path tracer flips accept/reject decisions.
// ---------------------------------------------------------------------------------------------

RT_HD unsigned long long d2ull(double d) {
#ifdef __CUDA_ARCH__
    return static_cast<unsigned long long>(__double_as_longlong(d));
#else
    unsigned long long u; memcpy(&u, &d, 8); return u;
#endif
}
RT_HD double ull2d(unsigned long long u) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double(static_cast<long long>(u));
#else
    double d; memcpy(&d, &u, 8); return d;
#endif
}

// f32::ln of ConstantMedium::hit (object.rs:562).  fdlibm-style: x = 2^k * m, m in [sqrt(1/2), sqrt(2)),
// log(m) from the s = f/(2+f) series.
RT_HD float ln_f32(float xf) {
#if defined(__CUDA_ARCH__) && RT_FAST_MATH
    return __logf(xf);  // tolerance build only
#endif
    if (xf != xf) return xf;
    if (xf < 0.f) return u2f(0x7fc00000u);
    if (xf == 0.f) return u2f(0xff800000u);
    if (xf == u2f(0x7f800000u)) return xf;
    double x = static_cast<double>(xf);
    unsigned long long bits = d2ull(x);
    uint32_t hx = static_cast<uint32_t>(bits >> 32);
    hx += 0x3ff00000u - 0x3fe6a09eu;
    int k = static_cast<int>(hx >> 20) - 0x3ff;
    hx = (hx & 0x000fffffu) + 0x3fe6a09eu;
    bits = (static_cast<unsigned long long>(hx) << 32) | (bits & 0xffffffffull);
    x = ull2d(bits);
    double f = x - 1.0;
    double hfsq = 0.5 * f * f;
    double s = f / (2.0 + f);
    double z = s * s;
    double w = z * z;
    double t1 = w * (3.999999999940941908e-01 + w * (2.222219843214978396e-01 + w * 1.531383769920937332e-01));
    double t2 = z * (6.666666666666735130e-01 +
                     w * (2.857142874366239149e-01 + w * (1.818357216161805012e-01 + w * 1.479819860511658591e-01)));
    double R = t2 + t1;
    double dk = static_cast<double>(k);
    double r = s * (hfsq + R) + dk * 1.90821492927058770002e-10 - hfsq + f + dk * 6.93147180369123816490e-01;
    return static_cast<float>(r);
}

// f32::powf(x, 5.) of schlick (material.rs:145): x^5 in double, rounded once.
RT_HD float pow5_f32(float xf) {
#if defined(__CUDA_ARCH__) && RT_FAST_MATH
    const float x2 = xf * xf;
    return x2 * x2 * xf;  // tolerance build only
#endif
    double d = static_cast<double>(xf);
    double d2 = d * d;
    double d4 = d2 * d2;
    return static_cast<float>(d4 * d);
}

// f32::sin of the checker texture (texture.rs:14): reduce by pi/2 in double, then the classic
// short odd/even polynomials in double.
RT_HD float sin_f32(float xf) {
#if defined(__CUDA_ARCH__) && RT_FAST_MATH
    return sinf(xf);  // tolerance build only (not __sinf: the checker's argument 10 * p is not small)
#endif
    if (xf != xf || fabsf(xf) == u2f(0x7f800000u)) return u2f(0x7fc00000u);
    if (fabsf(xf) < 0.000244140625f) return xf;
    double x = static_cast<double>(xf);
    double fn = (x * 6.36619772367581382433e-01 + 6755399441055744.0) - 6755399441055744.0;
    double y = (x - fn * 1.57079631090164184570e+00) - fn * 1.58932547735281966916e-08;
    double q = fn - 4.0 * floor(fn * 0.25);
    int n = static_cast<int>(q) & 3;
    double z = y * y;
    double w = z * z;
    double r;
    if (n & 1) {  // cosine kernel
        double c = ((1.0 + z * -0.499999997251031003120) + w * 0.0416666233237390631894) +
                   (w * z) * (-0.00138867637746099294692 + z * 0.0000243904487962774090654);
        r = (n == 1) ? c : -c;
    } else {      // sine kernel on +y (n == 0) or -y (n == 2); the polynomial is odd
        double yy = (n == 0) ? y : -y;
        double sgl = z * yy;
        r = (yy + sgl * (-0.166666666416265235595 + z * 0.0083333293858894631756)) +
            sgl * w * (-0.000198393348360966317347 + z * 0.0000027183114939898219064);
    }
    return static_cast<float>(r);
}

RT_HD float schlick(float cosine, float ref_idx) {  // material.rs:142-146
    float r0 = (1.f - ref_idx) / (1.f + ref_idx);
    r0 = r0 * r0;
    return r0 + (1.f - r0) * pow5_f32(1.f - cosine);
}

}  // namespace rtiow
