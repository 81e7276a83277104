// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product.
//
// Scalar math used by the CPU restatement of cbiffle/rtiow-rust's render path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
// build, link or call anything under oracle/.  The product (rtiow-rust_b200/) has its own,
// separately written device-side versions of everything in here.
//
// PARITY UNPINNED: the reference holds no golden vectors, known-answer tests or fixtures for this
// path (its only tests are two Vec3 indexing doctests, src/vec3.rs:221-228,270-277) and the Rust
// toolchain is absent, so nothing in here could be checked against reference output.  The KATs in
// tests/ are first-principles.
//
// Build flags that matter: -ffp-contract=off (Rust never contracts a*b+c into an FMA), no
// -ffast-math.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace oracle {

// ---------------------------------------------------------------------------------------------
// Vec3 — src/vec3.rs:13.  Every operator keeps the reference's association order.
// ---------------------------------------------------------------------------------------------
struct Vec3 {
    float x = 0.f, y = 0.f, z = 0.f;  // #[derive(Default)] -> (0,0,0)   vec3.rs:12
    Vec3() = default;
    Vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    float operator[](int axis) const { return axis == 0 ? x : (axis == 1 ? y : z); }  // vec3.rs:280-309
    float& operator[](int axis) { return axis == 0 ? x : (axis == 1 ? y : z); }
};

inline Vec3 splat(float v) { return Vec3(v, v, v); }                                    // vec3.rs:106-111
inline Vec3 operator*(Vec3 a, Vec3 b) { return Vec3(a.x * b.x, a.y * b.y, a.z * b.z); } // vec3.rs:115-122
inline Vec3 operator*(float s, Vec3 v) { return splat(s) * v; }                         // vec3.rs:125-132
inline Vec3 operator/(Vec3 a, Vec3 b) { return Vec3(a.x / b.x, a.y / b.y, a.z / b.z); } // vec3.rs:135-142
inline Vec3 operator/(Vec3 a, float s) { return Vec3(a.x / s, a.y / s, a.z / s); }      // vec3.rs:145-152
inline Vec3 operator+(Vec3 a, Vec3 b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); } // vec3.rs:155-162
inline Vec3 operator+(float s, Vec3 v) { return Vec3(s + v.x, s + v.y, s + v.z); }      // vec3.rs:165-172
inline Vec3 operator-(Vec3 a, Vec3 b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); } // vec3.rs:175-182
inline Vec3 operator-(Vec3 a) { return Vec3(-a.x, -a.y, -a.z); }                        // vec3.rs:185-192

// reduce(f) = f(f(x,y),z)   vec3.rs:100-102
inline float dot(Vec3 a, Vec3 b) {  // vec3.rs:43-46: zip_with(mul).reduce(add)
    Vec3 m = a * b;
    return (m.x + m.y) + m.z;
}
inline Vec3 cross(Vec3 a, Vec3 b) {  // vec3.rs:49-55
    return Vec3(a.y * b.z - a.z * b.y, -(a.x * b.z - a.z * b.x), a.x * b.y - a.y * b.x);
}
inline float length(Vec3 v) { return std::sqrt(dot(v, v)); }  // vec3.rs:59-61
inline Vec3 into_unit(Vec3 v) { return v / length(v); }       // vec3.rs:66-68

// f32::max / f32::min ignore a NaN operand (IEEE maxNum/minNum), as fmaxf/fminf do.
inline float rmax(float a, float b) { return std::fmax(a, b); }
inline float rmin(float a, float b) { return std::fmin(a, b); }

inline Vec3 reflect(Vec3 v, Vec3 n) {  // vec3.rs:313-315: `2. * v.dot(n) * n` == (2*dot)*n
    return v - (2.f * dot(v, n)) * n;
}

// vec3.rs:321-330.  Returns false for None.
inline bool refract(Vec3 v, Vec3 n, float ni_over_nt, Vec3& out) {
    Vec3 uv = into_unit(v);
    float dt = dot(uv, n);
    float discriminant = 1.0f - ni_over_nt * ni_over_nt * (1.f - dt * dt);
    if (discriminant > 0.f) {
        out = ni_over_nt * (uv - dt * n) - std::sqrt(discriminant) * n;
        return true;
    }
    return false;
}

// Rust `x as i32` for f32: saturating, NaN -> 0.
inline int32_t f32_as_i32(float x) {
    if (std::isnan(x)) return 0;
    if (x >= 2147483648.0f) return std::numeric_limits<int32_t>::max();
    if (x <= -2147483648.0f) return std::numeric_limits<int32_t>::min();
    return static_cast<int32_t>(x);
}

// ---------------------------------------------------------------------------------------------
// Transcendentals that run on the device in the product.  libm's logf/sinf/powf cannot be called
// from a kernel, and CUDA's own versions round differently, so the parity contract fixes ONE
// algorithm built from IEEE double +,-,*,/ only (bit-reproducible on any conforming target) and
// rounds once to f32 at the end.  This file and the product's device header each spell it out
// independently.  Against glibc these agree to <= 1 ulp (checked in tests/test_oracle_kat.py).
// ---------------------------------------------------------------------------------------------

// ln for material/medium code: object.rs:562 `rng().ln()`.
// Classic fdlibm e_log.c reduction, evaluated in double on the exactly-converted f32 input.
inline float log_f32(float xf) {
    if (std::isnan(xf)) return xf;
    if (xf < 0.f) return std::numeric_limits<float>::quiet_NaN();
    if (xf == 0.f) return -std::numeric_limits<float>::infinity();
    if (std::isinf(xf)) return xf;
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
                 Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
                 Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    double x = static_cast<double>(xf);  // exact; every nonzero f32 is a normal double
    uint64_t bits;
    std::memcpy(&bits, &x, 8);
    uint32_t hx = static_cast<uint32_t>(bits >> 32);
    int k = 0;
    hx += 0x3ff00000 - 0x3fe6a09e;
    k += static_cast<int>(hx >> 20) - 0x3ff;
    hx = (hx & 0x000fffff) + 0x3fe6a09e;
    bits = (static_cast<uint64_t>(hx) << 32) | (bits & 0xffffffffull);
    std::memcpy(&x, &bits, 8);  // x in [sqrt(2)/2, sqrt(2))
    double f = x - 1.0;
    double hfsq = 0.5 * f * f;
    double s = f / (2.0 + f);
    double z = s * s;
    double w = z * z;
    double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    double R = t2 + t1;
    double dk = static_cast<double>(k);
    double r = s * (hfsq + R) + dk * ln2_lo - hfsq + f + dk * ln2_hi;
    return static_cast<float>(r);
}

// powf(x, 5.) for schlick: material.rs:145.  x^5 in double (4 roundings at 2^-53) then one
// rounding to f32: equals the correctly rounded result except in ~2^-29 of cases.
inline float pow5_f32(float xf) {
    double d = static_cast<double>(xf);
    double d2 = d * d;
    double d4 = d2 * d2;
    double d5 = d4 * d;
    return static_cast<float>(d5);
}

// sin for the checker texture: texture.rs:14.  The well-known float-via-double scheme
// (Cody-Waite reduction by pi/2 in double, then degree-9/8 odd/even minimax kernels in double).
// For |x| >= 2^28*pi/2 the reduction loses accuracy (the contract keeps it deterministic rather
// than exact there; texture coordinates never get close).
namespace detail {
inline double sin_kernel(double x) {
    const double S1 = -0.166666666416265235595, S2 = 0.0083333293858894631756,
                 S3 = -0.000198393348360966317347, S4 = 0.0000027183114939898219064;
    double z = x * x;
    double w = z * z;
    double r = S3 + z * S4;
    double s = z * x;
    return (x + s * (S1 + z * S2)) + s * w * r;
}
inline double cos_kernel(double x) {
    const double C0 = -0.499999997251031003120, C1 = 0.0416666233237390631894,
                 C2 = -0.00138867637746099294692, C3 = 0.0000243904487962774090654;
    double z = x * x;
    double w = z * z;
    double r = C2 + z * C3;
    return ((1.0 + z * C0) + w * C1) + (w * z) * r;
}
}  // namespace detail

inline float sin_f32(float xf) {
    if (std::isnan(xf) || std::isinf(xf)) return std::numeric_limits<float>::quiet_NaN();
    const double invpio2 = 6.36619772367581382433e-01;
    const double pio2_1 = 1.57079631090164184570e+00;   // first 25 bits of pi/2
    const double pio2_1t = 1.58932547735281966916e-08;  // pi/2 - pio2_1
    const double toint = 6755399441055744.0;            // 1.5 * 2^52
    double x = static_cast<double>(xf);
    if (std::fabs(xf) < 0.000244140625f) return xf;  // |x| < 2^-12: sin(x) rounds to x
    double fn = (x * invpio2 + toint) - toint;       // nearest integer to x*2/pi
    double y = (x - fn * pio2_1) - fn * pio2_1t;
    // quadrant = fn mod 4, via exact double arithmetic (fn may exceed int32 for huge x)
    double q = fn - 4.0 * std::floor(fn * 0.25);
    int n = static_cast<int>(q);
    double r;
    switch (n & 3) {
        case 0: r = detail::sin_kernel(y); break;
        case 1: r = detail::cos_kernel(y); break;
        case 2: r = detail::sin_kernel(-y); break;
        default: r = -detail::cos_kernel(y); break;
    }
    return static_cast<float>(r);
}

// ---------------------------------------------------------------------------------------------
// Counter-based RNG: Philox4x32-10 (Salmon et al., SC'11).  Replaces the reference's sequential
// `&mut impl Rng` plumbing (lib.rs:24,60,367,384; object.rs:33; camera.rs:52; material.rs:55;
// vec3.rs:19-39,209-214), which is nondeterministic on the production path (thread_rng(),
// lib.rs:367) and inherently serial on the deterministic one (cast(), lib.rs:378).
// ---------------------------------------------------------------------------------------------
inline void philox4x32_10(const uint32_t key[2], const uint32_t ctr[4], uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = static_cast<uint64_t>(M0) * c0;
        uint64_t p1 = static_cast<uint64_t>(M1) * c2;
        uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = static_cast<uint32_t>(p1);
        uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = static_cast<uint32_t>(p0);
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// rand 0.6.5 `Standard` for f32: 24 high bits -> [0,1)   (un-vendored crate)
inline float u32_to_unit_f32(uint32_t w) { return static_cast<float>(w >> 8) * (1.0f / 16777216.0f); }
// rand 0.6.5 UniformFloat: 23 high bits as the mantissa of a float in [1,2)
inline float u32_to_f32_1_2(uint32_t w) {
    uint32_t b = 0x3F800000u | (w >> 9);
    float f;
    std::memcpy(&f, &b, 4);
    return f;
}

}  // namespace oracle
