// Shared device/host helpers for liboptk (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "optk.h"

namespace optk {

// ---------------------------------------------------------------------------
// error handling (host)
// ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define OPTK_CUDA(call)                                        \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return ::optk::cuda_fail(e__, #call); \
    } while (0)

// ---------------------------------------------------------------------------
// exact division of a 32-bit index by a runtime constant: q = umul64hi(n, M),
// M = floor((2^64 - 1) / d) + 1, exact for all n < 2^32 and 2 <= d < 2^32.
// ---------------------------------------------------------------------------
struct FastDiv {
    uint64_t magic;
    uint32_t divisor;
    uint32_t pad;
};

inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.divisor = d;
    f.pad = 0;
    f.magic = d > 1 ? (~0ull / d) + 1ull : 0ull;
    return f;
}

__device__ __forceinline__ void divmod(uint32_t n, const FastDiv& f, uint32_t& q, uint32_t& r) {
    q = (uint32_t)__umul64hi((uint64_t)n, f.magic);
    r = n - q * f.divisor;
}

// ---------------------------------------------------------------------------
// fp64 helpers.  "exact" variants use the _rn intrinsics, which nvcc never
// contracts into FMAs: they reproduce NumPy's operation-by-operation rounding
// for the edge-sensitive comparisons (aperture edges, bin edges).
// ---------------------------------------------------------------------------
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }

// numpy.sign: -1, 0, +1 (NaN -> NaN)
__device__ __forceinline__ double sign0(double x) {
    return x != x ? x : (double)((x > 0.0) - (x < 0.0));
}

__device__ __forceinline__ double sq(double x) { return x * x; }

// ---------------------------------------------------------------------------
// fp64 reciprocal / division / square root without the IEEE slow paths.
// B200 has no fp64 divide or sqrt unit: `a / b` and `sqrt(x)` compile to a
// MUFU seed, Newton steps, a residual correction AND exponent-range fix-ups
// (20-25 instructions each).  For operands whose exponent lies in
// [2^-511, 2^512) the fix-ups are dead weight: seed (20+ bits) -> two Newton
// steps -> one residual correction gives a faithfully rounded result (error
// < 1 ulp) in 8-11 instructions.  Anything else (zero, subnormal, huge, inf,
// NaN, negative under a root) takes the IEEE path, so NaN / inf propagation is
// exactly that of the plain operators.  Used where 1 ulp is irrelevant against
// the 1e-9 parity tolerance; edge-sensitive comparisons keep the exact forms.
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool exponent_mid(double x) {
    const unsigned e = (unsigned)__double2hiint(x) & 0x7ff00000u;
    return (e - 0x20000000u) < 0x40000000u;  // biased exponent in [0x200, 0x600)
}

__device__ __forceinline__ double rcp_newton(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

__device__ __forceinline__ double frcp(double x) {
    if (!exponent_mid(x)) return 1.0 / x;
    return rcp_newton(x);
}

__device__ __forceinline__ double fdiv(double a, double b) {
    // a == 0 is fine on the fast path; a non-finite or huge is not
    const unsigned ea = (unsigned)__double2hiint(a) & 0x7ff00000u;
    if (!exponent_mid(b) || ea >= 0x60000000u) return a / b;
    const double y = rcp_newton(b);
    const double q = a * y;
    return fma(fma(-b, q, a), y, q);
}

__device__ __forceinline__ double fsqrt(double x) {
    if (!exponent_mid(x) || x < 0.0) return sqrt(x);
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    return fma(fma(-g, g, x), h, g);
}

// 1 / sqrt(x)
__device__ __forceinline__ double frsqrt(double x) {
    if (!exponent_mid(x) || x < 0.0) return 1.0 / sqrt(x);
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x * y, y, 1.0);  // 1 - x y^2
    y = fma(0.5 * y, e, y);
    e = fma(-x * y, y, 1.0);
    y = fma(y * fma(0.375, e, 0.5), e, y);  // second step with the e^2 term
    return y;
}

// x -> R x + t
__device__ __forceinline__ void affine_forward(const optk_affine_t& a, double& x, double& y, double& z,
                                               bool is_direction) {
    double rx = a.r[0] * x + a.r[1] * y + a.r[2] * z;
    double ry = a.r[3] * x + a.r[4] * y + a.r[5] * z;
    double rz = a.r[6] * x + a.r[7] * y + a.r[8] * z;
    if (!is_direction) {
        rx += a.t[0];
        ry += a.t[1];
        rz += a.t[2];
    }
    x = rx; y = ry; z = rz;
}

// x -> R^T (x - t)
__device__ __forceinline__ void affine_inverse(const optk_affine_t& a, double& x, double& y, double& z,
                                               bool is_direction) {
    if (!is_direction) {
        x -= a.t[0];
        y -= a.t[1];
        z -= a.t[2];
    }
    double rx = a.r[0] * x + a.r[3] * y + a.r[6] * z;
    double ry = a.r[1] * x + a.r[4] * y + a.r[7] * z;
    double rz = a.r[2] * x + a.r[5] * y + a.r[8] * z;
    x = rx; y = ry; z = rz;
}

}  // namespace optk
