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
// fp64 reciprocal / division / square root, branch-free.
// B200 has no fp64 divide or sqrt unit: `a / b` and `sqrt(x)` compile to a MUFU seed,
// Newton steps, a residual correction AND exponent-range fix-ups behind a branch and a
// call (20-25 instructions, and the branch stops the scheduler from interleaving the
// two rays a thread carries).  Here: seed (20+ bits) -> two Newton steps -> one residual
// correction = faithfully rounded (error < 1 ulp, >99.9 % correctly rounded) in 9-12
// straight-line instructions.  Zero / inf / NaN operands are repaired with selects from the
// raw seed, which already has the IEEE special values (rcp(0) = inf, rcp(inf) = 0,
// rsqrt(0) = inf, rsqrt(<0) = NaN), so NaN / inf propagation matches the plain operators.
// Subnormal operands are flushed to zero and intermediate overflow beyond 1e300 is not
// rescued (lengths are millimetres).  Used where 1 ulp is irrelevant against the 1e-9
// parity tolerance; edge-sensitive comparisons keep exact arithmetic.
// ---------------------------------------------------------------------------
#define OPTK_INF __longlong_as_double(0x7ff0000000000000LL)
#define OPTK_NAN __longlong_as_double(0x7ff8000000000000LL)

__device__ __forceinline__ double rcp_seed(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}

__device__ __forceinline__ double rsqrt_seed(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}

__device__ __forceinline__ double frcp(double x) {
    const double y0 = rcp_seed(x);
    double e = fma(-x, y0, 1.0);
    double y = fma(y0, e, y0);
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    return (fabs(y) < OPTK_INF) ? y : y0;  // x = 0, inf, NaN: the seed is the answer
}

__device__ __forceinline__ double fdiv(double a, double b) {
    const double y0 = rcp_seed(b);
    double e = fma(-b, y0, 1.0);
    double y = fma(y0, e, y0);
    e = fma(-b, y, 1.0);
    y = fma(y, e, y);
    const double q0 = a * y;
    const double q = fma(fma(-b, q0, a), y, q0);
    return (fabs(q) < OPTK_INF) ? q : a * y0;  // b = 0 / inf, a = inf, NaN: IEEE values from the seed
}

__device__ __forceinline__ double fsqrt(double x) {
    const double y0 = rsqrt_seed(x);
    double g = x * y0, h = 0.5 * y0;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    g = fma(fma(-g, g, x), h, g);
    // x = 0 (or subnormal), +inf, negative, NaN
    const double special = (x >= 0.0) ? ((x == OPTK_INF) ? x : 0.0) : OPTK_NAN;
    return (fabs(g) < OPTK_INF) ? g : special;
}

// 1 / sqrt(x)
__device__ __forceinline__ double frsqrt(double x) {
    const double y0 = rsqrt_seed(x);
    double e = fma(-x * y0, y0, 1.0);  // 1 - x y^2
    double y = fma(0.5 * y0, e, y0);
    e = fma(-x * y, y, 1.0);
    y = fma(y * fma(0.375, e, 0.5), e, y);  // second step with the e^2 term
    return (fabs(y) < OPTK_INF) ? y : y0;
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
