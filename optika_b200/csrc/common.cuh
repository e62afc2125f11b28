// Shared device/host helpers for liboptk (sm_100a).
#pragma once
#ifdef __CUDACC_RTC__
// run-time compilation of a kernel specialised for one system (jit.cu): no host headers
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
typedef unsigned long uintptr_t;
#define NAN __longlong_as_double(0x7ff8000000000000LL)
#define INFINITY __longlong_as_double(0x7ff0000000000000LL)
#else
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#endif
#include "optk.h"

namespace optk {

#ifndef __CUDACC_RTC__
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
#endif

// ---------------------------------------------------------------------------
// exact division of a 32-bit index by a runtime constant: q = umul64hi(n, M),
// M = floor((2^64 - 1) / d) + 1, exact for all n < 2^32 and 2 <= d < 2^32 (d = 1: M = 0 marks q = n).
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
    q = f.magic ? (uint32_t)__umul64hi((uint64_t)n, f.magic) : n;  // magic 0: d = 1
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

// Dot products and sums of squares with the multiply-adds written out.  `a * b + c * d + e * f` leaves the
// choice of which products are fused to the compiler, and the choice can differ between two copies of the same
// expression (nvcc / NVRTC, a tail the optimiser duplicated for a rare branch): the table-driven and the
// run-time compiled kernels are held to bit-identical results, so the hot paths say what they mean.
__device__ __forceinline__ double dot3(double ax, double ay, double az, double bx, double by, double bz) {
    return fma(ax, bx, fma(ay, by, az * bz));
}
__device__ __forceinline__ double norm2_3(double x, double y, double z) { return fma(x, x, fma(y, y, z * z)); }
__device__ __forceinline__ double norm2_2(double x, double y) { return fma(x, x, y * y); }

// ---------------------------------------------------------------------------
// fp64 reciprocal / division / square root, branch-free.
// B200 has no fp64 divide or sqrt unit: `a / b` and `sqrt(x)` compile to a MUFU seed,
// Newton steps, a residual correction AND exponent-range fix-ups behind a branch and a
// call (20-25 instructions, and the branch stops the scheduler from interleaving the
// two rays a thread carries).  Here: seed (20 bits) -> one Newton step -> one residual
// correction (or one third-order step) = error < 1 ulp in 6-9 straight-line instructions.  Zero / inf / NaN operands are repaired with selects from the
// raw seed, which already has the IEEE special values (rcp(0) = inf, rcp(inf) = 0,
// rsqrt(0) = inf, rsqrt(<0) = NaN), so NaN / inf propagation matches the plain operators.
// Subnormal operands are flushed to zero and intermediate overflow beyond 1e300 is not
// rescued (lengths are millimetres).  Used where 1 ulp is irrelevant against the 1e-9
// parity tolerance; edge-sensitive comparisons keep exact arithmetic.
// ---------------------------------------------------------------------------
#define OPTK_INF __longlong_as_double(0x7ff0000000000000LL)
#define OPTK_NAN __longlong_as_double(0x7ff8000000000000LL)

// inf or NaN?  Decided on the exponent bits with two integer instructions (ALU pipe) instead of a
// DSETP: the trace kernels are bound by the FP64 pipe, where every compare costs an issue slot.
__device__ __forceinline__ bool not_finite(double y) {
    return ((unsigned)__double2hiint(y) & 0x7ff00000u) == 0x7ff00000u;
}

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

// The seeds carry the leading 20 bits of the mantissa (MUFU.RCP64H / RSQ64H read and write the high word
// only): relative error e0 <= 2^-19.  One third-order step takes that to e0^3 < 2^-57, below the rounding of
// the last operation; a second step (round 1) only recomputed digits that were already right.
// tests/test_gpu_math.py measures all four against correctly rounded results (<= 1 ulp) through
// optk_debug_math.
__device__ __forceinline__ double frcp_raw(double x) {
    const double y0 = rcp_seed(x);
    const double e = fma(-x, y0, 1.0);
    return fma(e, fma(y0, e, y0), y0);  // y0 (1 + e + e^2)
}

__device__ __forceinline__ double frcp(double x) {
    const double y0 = rcp_seed(x);
    const double e = fma(-x, y0, 1.0);
    const double y = fma(e, fma(y0, e, y0), y0);
    return not_finite(y) ? y0 : y;  // x = 0, inf, NaN: the seed is the answer
}

__device__ __forceinline__ double fdiv(double a, double b) {
    const double y0 = rcp_seed(b);
    const double e = fma(-b, y0, 1.0);
    const double y = fma(y0, e, y0);  // 1 / b to 2^-38: the residual step below squares that
    const double q0 = a * y;
    double q = fma(fma(-b, q0, a), y, q0);
    // b = 0 / inf, a = inf, NaN: the IEEE value comes from the seed.  One predicated multiply
    // instead of a multiply and a two-word select.
    asm("{\n"
        ".reg .pred p;\n"
        ".reg .b32 lo, hi;\n"
        "mov.b64 {lo, hi}, %0;\n"
        "and.b32 hi, hi, 0x7ff00000;\n"
        "setp.eq.u32 p, hi, 0x7ff00000;\n"
        "@p mul.f64 %0, %1, %2;\n"
        "}"
        : "+d"(q)
        : "d"(a), "d"(y0));
    return q;
}

// a / b to 2^-38 (one Newton step on the seed, no residual correction), NaN instead of inf for b = 0: for the
// steps of an iteration that corrects itself -- a step of size s lands within 4e-12 s of where the exact
// quotient would, and the next step (or the convergence test, which looks at s) absorbs that.
__device__ __forceinline__ double fdiv_newton(double a, double b) {
    const double y0 = rcp_seed(b);
    const double e = fma(-b, y0, 1.0);
    return a * fma(y0, e, y0);
}

// a / b for a divisor that is never infinite (a direction component, a squared length): no residual step
// (error <= 1.5 ulp), and the special values come out of the arithmetic instead of a repair -- the seed of
// b = +-0 is +-inf, clamped here to +-2^1023 with two integer min, so that 1 - b y0 = 1 instead of NaN and
// y0 (1 + e + e^2) overflows to the IEEE +-inf; a = +-inf, NaN in either operand and 0 / 0 follow by themselves.
// (b = +-inf would give NaN instead of 0.)  7 instructions, two of them off the FP64 pipe, against 11.
__device__ __forceinline__ double fdiv_finite(double a, double b) {
    double y0;
    asm("{\n"
        ".reg .b32 lo, hi;\n"
        ".reg .f64 s;\n"
        "rcp.approx.ftz.f64 s, %1;\n"
        "mov.b64 {lo, hi}, s;\n"
        "min.s32 hi, hi, 0x7fe00000;\n"
        "min.u32 hi, hi, 0xffe00000;\n"
        "mov.b64 %0, {lo, hi};\n"
        "}"
        : "=d"(y0)
        : "d"(b));
    const double e = fma(-b, y0, 1.0);
    return a * fma(e, fma(y0, e, y0), y0);
}

__device__ __forceinline__ double fsqrt(double x) {
    const double y0 = rsqrt_seed(x);
    double g = x * y0, h = 0.5 * y0;
    const double r = fma(-g, h, 0.5);
    g = fma(g, r, g);  // sqrt(x) and 1 / (2 sqrt(x)) to 2^-38
    h = fma(h, r, h);
    g = fma(fma(-g, g, x), h, g);
    // x = 0 (seed inf), +inf (seed 0), negative or NaN (seed NaN) all leave g = NaN.  The
    // first two have sqrt(x) = x and are exactly the non-negative fixed points of x + x = x;
    // for the others NaN is the answer.  (A subnormal radicand is flushed by the seed and also
    // yields NaN; radicands here are sums of squares of millimetre-scale quantities.)
    // (decided on the bit patterns, off the FP64 pipe: g is NaN in exactly those cases, and
    // "x >= 0 or x is NaN with a clear sign" is one unsigned compare of the high word, -0 included)
    return (not_finite(g) && (unsigned)__double2hiint(x) <= 0x80000000u) ? x : g;
}

// 1 / sqrt(x) without the repair of x = 0 / inf (both give NaN here): for radicands that are >= 1 by
// construction (1 + slopes^2) or whose zero is a NaN of the caller anyway (the toroid's domain boundary).
__device__ __forceinline__ double frsqrt_raw(double x) {
    const double y0 = rsqrt_seed(x);
    const double e = fma(-x * y0, y0, 1.0);  // 1 - x y0^2
    return fma(y0 * e, fma(0.375, e, 0.5), y0);  // y0 (1 + e / 2 + 3 e^2 / 8)
}

// 1 / sqrt(x)
__device__ __forceinline__ double frsqrt(double x) {
    const double y0 = rsqrt_seed(x);
    const double e = fma(-x * y0, y0, 1.0);
    const double y = fma(y0 * e, fma(0.375, e, 0.5), y0);
    return not_finite(y) ? y0 : y;
}

// ---------------------------------------------------------------------------
// exp and sincos for the multilayer kernel: straight-line, ~20 / ~32 instructions instead of
// the library's 40-75 with their slow-path calls.  exp: |error| < 2 ulp for |x| <= 700, inputs
// beyond are clamped (exp(700) ~ 1e304 is already past every guard in the model), NaN
// propagates.  sincos: three-term Cody-Waite reduction, absolute error < 1e-15 for
// |x| < 1e5; larger arguments (a layer thousands of wavelengths thick) take the library path.
// ---------------------------------------------------------------------------
static __constant__ double kExp[16] = {
    1.6059043836821613e-10,  // 1/13!
    2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07, 2.7557319223985893e-06,
    2.48015873015873e-05, 1.984126984126984e-04, 1.388888888888889e-03, 8.333333333333333e-03,
    4.1666666666666664e-02, 1.6666666666666666e-01, 0.5, 1.0, 1.0,
    -6.93147180369123816490e-01, -1.90821492927058770002e-10,  // -ln2 hi, lo
};
static __constant__ double kSin[6] = {
    1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,
    -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01,
};
static __constant__ double kCos[6] = {
    -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,
    2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02,
};
static __constant__ double kPio2[3] = {
    -1.57079632673412561417e+00, -6.07710050650619224932e-11, -2.02226624879595063154e-21,  // pi/2 in three parts
};

__device__ __forceinline__ double fexp(double x) {
    const double xc = fmin(fmax(x, -700.0), 700.0);
    const double shifter = 6755399441055744.0;  // 1.5 * 2^52: rounds to integer in the low word
    const double kd = fma(xc, 1.4426950408889634, shifter);
    const int k = __double2loint(kd);
    const double kf = kd - shifter;
    double r = fma(kf, kExp[14], xc);  // -ln2 hi
    r = fma(kf, kExp[15], r);          // -ln2 lo
    // Coefficients come from the constant bank (an operand of the DFMA itself).  Estrin's
    // scheme: 16 operations in a dependency chain of depth 5 instead of Horner's 13 dependent
    // FMAs -- the multilayer kernel is bound by fp64 latency ("wait" stalls), not by the pipe.
    // a_j (degree j) = kExp[13 - j].
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double q0 = fma(kExp[12], r, kExp[13]), q1 = fma(kExp[10], r, kExp[11]);
    const double q2 = fma(kExp[8], r, kExp[9]), q3 = fma(kExp[6], r, kExp[7]);
    const double q4 = fma(kExp[4], r, kExp[5]), q5 = fma(kExp[2], r, kExp[3]);
    const double q6 = fma(kExp[0], r, kExp[1]);
    const double u0 = fma(q1, r2, q0), u1 = fma(q3, r2, q2), u2 = fma(q5, r2, q4);
    const double t0 = fma(u1, r4, u0), t1 = fma(q6, r4, u2);
    double p = fma(t1, r8, t0);
    // scale by 2^k in two halves so that k in [-1010, 1010] never leaves the normal range midway
    const int k1 = k >> 1, k2 = k - k1;
    const double s1 = __hiloint2double((k1 + 1023) << 20, 0), s2 = __hiloint2double((k2 + 1023) << 20, 0);
    const double y = p * s1 * s2;
    return (x != x) ? x : y;
}

__device__ __forceinline__ void fsincos(double x, double* s, double* c) {
    if (!(fabs(x) < 1.0e5)) {
        sincos(x, s, c);
        return;
    }
    const double shifter = 6755399441055744.0;
    const double kd = fma(x, 6.36619772367581382433e-01, shifter);  // x * 2/pi
    const int q = __double2loint(kd);
    const double kf = kd - shifter;
    double r = fma(kf, kPio2[0], x);  // pi/2, first 33 bits
    r = fma(kf, kPio2[1], r);         // next 33 bits
    r = fma(kf, kPio2[2], r);         // tail
    const double r2 = r * r;
    // sin(r), cos(r) on [-pi/4, pi/4]
    double ps = kSin[0], pc = kCos[0];
#pragma unroll
    for (int i = 1; i < 6; ++i) {
        ps = fma(ps, r2, kSin[i]);
        pc = fma(pc, r2, kCos[i]);
    }
    const double sr = fma(ps * r2, r, r);
    const double cr = fma(pc * r2, r2, fma(-0.5, r2, 1.0));
    // quadrant: q mod 4 = 0: (s, c); 1: (c, -s); 2: (-s, -c); 3: (-c, s)
    const bool swap = q & 1;
    const double ss = swap ? cr : sr, cc = swap ? sr : cr;
    *s = (q & 2) ? -ss : ss;
    *c = ((q + 1) & 2) ? -cc : cc;
}

// x -> R x + t
__device__ __forceinline__ void affine_forward(const optk_affine_t& a, double& x, double& y, double& z,
                                               bool is_direction) {
    double rx, ry, rz;
    // (explicit multiply-adds: nvcc and NVRTC then round the same way, and the table-driven and the run-time
    // compiled kernels stay bit-identical)
    if (is_direction) {
        rx = fma(a.r[0], x, fma(a.r[1], y, a.r[2] * z));
        ry = fma(a.r[3], x, fma(a.r[4], y, a.r[5] * z));
        rz = fma(a.r[6], x, fma(a.r[7], y, a.r[8] * z));
    } else {  // the offset rides in the innermost multiply-add: three instructions per component, not four
        rx = fma(a.r[0], x, fma(a.r[1], y, fma(a.r[2], z, a.t[0])));
        ry = fma(a.r[3], x, fma(a.r[4], y, fma(a.r[5], z, a.t[1])));
        rz = fma(a.r[6], x, fma(a.r[7], y, fma(a.r[8], z, a.t[2])));
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
    double rx = fma(a.r[0], x, fma(a.r[3], y, a.r[6] * z));
    double ry = fma(a.r[1], x, fma(a.r[4], y, a.r[7] * z));
    double rz = fma(a.r[2], x, fma(a.r[5], y, a.r[8] * z));
    x = rx; y = ry; z = rz;
}

}  // namespace optk
