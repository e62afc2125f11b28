// Kernel 3: transfer-matrix (Yeh) efficiency of a multilayer stack (sm_100a, fp64).
//
// One thread evaluates one point of the (wavelength x angle x configuration)
// grid.  The layer table is staged in shared memory; the complex 2x2 chain for
// both polarisations stays in registers.  Restates
// optika/materials/_multilayers.py:187-237, 485-532, _layers.py:243-277, 487-499,
// 625-645, matrices.py:127-161, 241-246 and profiles.py:103-126.
#include "common.cuh"
#include "params.cuh"
#include <cstdlib>

namespace optk {

struct cplx {
    double re, im;
};

__device__ __forceinline__ cplx C(double re, double im = 0.0) { return cplx{re, im}; }
__device__ __forceinline__ cplx operator+(cplx a, cplx b) { return C(a.re + b.re, a.im + b.im); }
__device__ __forceinline__ cplx operator-(cplx a, cplx b) { return C(a.re - b.re, a.im - b.im); }
__device__ __forceinline__ cplx operator*(cplx a, cplx b) {
    return C(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
__device__ __forceinline__ cplx operator*(double a, cplx b) { return C(a * b.re, a * b.im); }
__device__ __forceinline__ cplx conj(cplx a) { return C(a.re, -a.im); }
__device__ __forceinline__ double norm2(cplx a) { return a.re * a.re + a.im * a.im; }
__device__ __forceinline__ cplx cinv(cplx a) {
    const double s = frcp(norm2(a));
    return C(a.re * s, -a.im * s);
}
__device__ __forceinline__ cplx operator/(cplx a, cplx b) { return a * cinv(b); }
__device__ __forceinline__ cplx csel(bool c, cplx a, cplx b) { return C(c ? a.re : b.re, c ? a.im : b.im); }

// principal square root with the C99 / numpy branch conventions, branch-free
__device__ __forceinline__ cplx csqrt(cplx z) {
    const double m = fsqrt(norm2(z));
    const double x = 0.5 * (m + fabs(z.re));  // t^2 with t = sqrt((|z| + |Re z|) / 2)
    const double it = frsqrt(x);              // 1 / t  (inf when z == 0)
    const double t = x * it;
    const double o = 0.5 * z.im * it;         // Im z / (2 t)
    // Re z >= 0: (t, Im z / 2t);  Re z < 0: (|Im z| / 2t, sign(Im z) t);  z == 0: (0, Im z)
    const bool right = z.re >= 0.0;
    cplx r = C(right ? t : fabs(o), right ? o : copysign(t, z.im));
    if (x == 0.0) r = C(0.0, z.im);
    return r;
}

struct Mat2 {
    cplx a, b, c, d;  // [[a, b], [c, d]]
};

__device__ __forceinline__ Mat2 identity2() { return Mat2{C(1.0), C(0.0), C(0.0), C(1.0)}; }

__device__ __forceinline__ Mat2 matmul(const Mat2& x, const Mat2& y) {
    return Mat2{x.a * y.a + x.b * y.c, x.a * y.b + x.b * y.d, x.c * y.a + x.d * y.c, x.c * y.b + x.d * y.d};
}

// The transfer matrix of one polarisation, kept UNNORMALISED:
//   true M = m * u / d     with  d = prod(2 q_i)  (the 1 / t_ij factors of matrices.py:156-160)
//                          and   u = prod(exp(-i beta_j))  (factored out of U, matrices.py:241-246).
// r = M21 / M11 does not see the scalars; t = 1 / M11 = d / (m.a * u).
struct PolChain {
    Mat2 m;
    cplx d;
};

struct Chain {
    PolChain s, p;
    cplx u;  // shared by both polarisations
};

__device__ __forceinline__ Chain chain_identity() {
    return Chain{PolChain{identity2(), C(1.0)}, PolChain{identity2(), C(1.0)}, C(1.0)};
}

__device__ __forceinline__ Chain chain_mul(const Chain& x, const Chain& y) {
    return Chain{PolChain{matmul(x.s.m, y.s.m), x.s.d * y.s.d}, PolChain{matmul(x.p.m, y.p.m), x.p.d * y.p.d},
                 x.u * y.u};
}

// n-fold product by square-and-multiply (Cartesian2dMatrixArray.power, _layers.py:643)
__device__ __forceinline__ Chain chain_pow(Chain x, int n) {
    Chain r = chain_identity();
    while (n > 0) {
        if (n & 1) r = chain_mul(r, x);
        n >>= 1;
        if (n) x = chain_mul(x, x);
    }
    return r;
}

// profiles.py:103-126 + the four _derivative_fourier_transform bodies
__device__ __forceinline__ double interface_factor(int kind, double width, double s) {
    const double pi = 3.141592653589793;
    switch (kind) {
        case 1: {
            const double x = s * width;
            return fexp(-(x * x) / 2.0);
        }
        case 2: {
            const double x = s * width;
            return frcp(1.0 + (x * x) / 2.0);
        }
        case 3: {
            const double x = 1.7320508075688772 * width * s;
            return sin(x) / x;
        }
        case 4: {
            const double a = pi / (pi * pi - 8.0);
            const double x = a * width * s;
            const double x1 = x - pi / 2.0, x2 = x + pi / 2.0;
            return pi * (sin(x1) / x1 + sin(x2) / x2) / 4.0;
        }
    }
    return 1.0;
}

// The medium the light is in: index, cosine of the propagation angle, q_s = n cos, q_p = conj(cos) / n.
struct Medium {
    cplx n, dir, q_s, q_p;
};

// M <- M * [[a, b], [b, a]] * diag(1, w)
__device__ __forceinline__ void pol_step(PolChain& c, cplx a, cplx b, cplx w, cplx two_q) {
    const Mat2 m = c.m;
    c.m.a = m.a * a + m.b * b;
    c.m.b = (m.a * b + m.b * a) * w;
    c.m.c = m.c * a + m.d * b;
    c.m.d = (m.c * b + m.d * a) * w;
    c.d = c.d * two_q;
}

// Layer.transfer (_layers.py:229-277) for both polarisations, accumulated into `acc`.
//   snells_law_scalar (_snells_law.py:34-38): n sin(theta) is invariant through the stack, so
//     cos(theta_j) = sqrt(1 - K2 / n_j^2) with K2 = n_0^2 (1 - cos(theta_0)^2) -- the same value the
//     reference reaches through sqrt(1 - cos^2) of the previous layer, without that root;
//   matrices.refraction (matrices.py:127-161): W = [[1, r], [r, 1]] / t with
//     r = (q_i - q_j) / (q_i + q_j) * roughness, t = 2 q_i / (q_i + q_j)
//     = [[q_i + q_j, (q_i - q_j) roughness], [same, q_i + q_j]] / (2 q_i);
//   matrices.propagation (matrices.py:241-246): U = exp(-i beta) diag(1, exp(2 i beta)),
//     beta = 2 pi h n cos(theta) / lambda = (2 pi h / lambda) q_s;
//   the |U_00| < 1e10 guard and the identity substitution (_layers.py:261-277).
// `is_substrate` forces the thickness to zero (_multilayers.py:190-193).
__device__ __forceinline__ void layer_step(const LayerDev& L, const unsigned* idx, int n_axes, double wavelength,
                                           cplx k2, bool is_substrate, Medium& med, bool& where, Chain& acc) {
    // every array of a layer usually varies along ONE grid axis (n_j with wavelength, thickness
    // with configuration): one multiply instead of a loop over the axes
    auto offset = [&](int axis, const long long* stride) -> long long {
        if (axis == -1) return 0;
        if (axis >= 0) {
            const unsigned i = axis == 0 ? idx[0] : (axis == 1 ? idx[1] : (axis == 2 ? idx[2] : idx[3]));
            const long long st = axis == 0 ? stride[0] : (axis == 1 ? stride[1] : (axis == 2 ? stride[2] : stride[3]));
            return (long long)i * st;
        }
        long long o = 0;
        for (int a = 0; a < n_axes; ++a) o += (long long)idx[a] * stride[a];
        return o;
    };
    const long long on = offset(L.n_axis, L.n_stride);
    const long long ot = offset(L.t_axis, L.t_stride);
    const long long ow = offset(L.w_axis, L.w_stride);
    const cplx n_j = C(__ldg(L.n_re + on), L.n_im ? __ldg(L.n_im + on) : 0.0);
    const double h = (is_substrate || !L.thickness) ? 0.0 : __ldg(L.thickness + ot);

    const cplx inv_n = cinv(n_j);
    const cplx dir_j = csqrt(C(1.0) - k2 * (inv_n * inv_n));
    const cplx q_s_j = dir_j * n_j;
    const cplx q_p_j = conj(dir_j) * inv_n;

    double rough_s = 1.0, rough_p = 1.0;
    if (L.profile_kind) {
        // s = Re(4 pi n_i direction_i / wavelength); direction_i is conjugated for p
        const double width = L.width ? __ldg(L.width + ow) : 0.0;
        const double k = 4.0 * 3.141592653589793 / wavelength;
        rough_s = interface_factor(L.profile_kind, width, k * med.q_s.re);
        rough_p = interface_factor(L.profile_kind, width, k * (med.n * conj(med.dir)).re);
    }

    // propagation through this layer
    const cplx beta = (2.0 * 3.141592653589793 * h / wavelength) * q_s_j;
    const double growth = fexp(beta.im);  // |exp(-i beta)|
    const bool where_new = where && (growth < 1e10);  // _layers.py:271-272 (NaN -> false)
    double sn1, cs1;
    fsincos(beta.re, &sn1, &cs1);
    const double decay = frcp(growth * growth);  // |exp(2 i beta)| = exp(-2 Im beta)
    // exp(2 i beta) from the double-angle identities: one sincos per layer
    const cplx w = where_new ? C(decay * (cs1 * cs1 - sn1 * sn1), decay * (2.0 * sn1 * cs1)) : C(1.0);
    const cplx u0 = where_new ? C(growth * cs1, -growth * sn1) : C(1.0);  // exp(-i beta)

    // refraction = where(where, refraction, identity) uses the mask BEFORE this layer (_layers.py:261-262)
    {
        const cplx a = csel(where, med.q_s + q_s_j, C(1.0));
        const cplx b = csel(where, rough_s * (med.q_s - q_s_j), C(0.0));
        pol_step(acc.s, a, b, w, csel(where, 2.0 * med.q_s, C(1.0)));
    }
    {
        const cplx a = csel(where, med.q_p + q_p_j, C(1.0));
        const cplx b = csel(where, rough_p * (med.q_p - q_p_j), C(0.0));
        pol_step(acc.p, a, b, w, csel(where, 2.0 * med.q_p, C(1.0)));
    }
    acc.u = acc.u * u0;
    where = where_new;
    med.n = n_j;
    med.dir = dir_j;
    med.q_s = q_s_j;
    med.q_p = q_p_j;
}

// PERIODIC: some segment repeats (needs a second chain and the matrix power); the explicit-stack
// kernel is compiled without it so that it fits 4 CTAs of 128 threads per SM.
template <bool PERIODIC, int MINB>
__global__ void __launch_bounds__(128, MINB) multilayer_kernel(const __grid_constant__ MultilayerParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LayerDev* layers = reinterpret_cast<LayerDev*>(smem_raw);
    {
        // stage the layer table in shared memory
        const int n_words = P.n_layers * (int)(sizeof(LayerDev) / sizeof(long long));
        const long long* src = reinterpret_cast<const long long*>(P.layers);
        long long* dst = reinterpret_cast<long long*>(layers);
        for (int k = threadIdx.x; k < n_words; k += blockDim.x) dst[k] = src[k];
    }
    __syncthreads();

    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= P.n_eval) return;

    unsigned idx[OPTK_ML_MAX_AXES] = {0, 0, 0, 0};
    {
        uint32_t rem = (uint32_t)e;
        for (int a = P.in.n_axes - 1; a >= 0; --a) {
            if (a == 0) {
                idx[a] = rem;
            } else {
                uint32_t q, r;
                divmod(rem, P.div[a], q, r);
                idx[a] = r;
                rem = q;
            }
        }
    }
    long long ow = 0, od = 0, on = 0;
    for (int a = 0; a < P.in.n_axes; ++a) {
        ow += (long long)idx[a] * P.in.wavelength_stride[a];
        od += (long long)idx[a] * P.in.direction_stride[a];
        on += (long long)idx[a] * P.in.n_stride[a];
    }
    const double wavelength = __ldg(P.in.wavelength + ow);
    const cplx dir0 = C(__ldg(P.in.direction_re + od), P.in.direction_im ? __ldg(P.in.direction_im + od) : 0.0);
    const cplx n0 = C(__ldg(P.in.n_re + on), P.in.n_im ? __ldg(P.in.n_im + on) : 0.0);

    Medium med;
    med.n = n0;
    med.dir = dir0;
    med.q_s = dir0 * n0;
    med.q_p = conj(dir0) / n0;
    const cplx q_amb_s = med.q_s, q_amb_p = med.q_p;
    const cplx k2 = (n0 * n0) * (C(1.0) - dir0 * dir0);  // (n sin(theta))^2, invariant through the stack
    bool where = true;

    Chain total = chain_identity();
    const int n_axes = P.in.n_axes;

    for (int g = 0; g < P.n_segments; ++g) {
        const optk_ml_segment_t seg = P.segments[g];
        if (seg.repeat == 1) {
            // LayerSequence.transfer, _layers.py:487-499
            for (int j = seg.first; j < seg.first + seg.count; ++j)
                layer_step(layers[j], idx, n_axes, wavelength, k2, false, med, where, total);
        } else if (PERIODIC && seg.repeat > 1) {
            // PeriodicLayerSequence.transfer, _layers.py:625-645: the first period
            // explicitly, the second period raised to the power (repeat - 1)
            for (int j = seg.first; j < seg.first + seg.count; ++j)
                layer_step(layers[j], idx, n_axes, wavelength, k2, false, med, where, total);
            Chain period = chain_identity();
            for (int j = seg.first; j < seg.first + seg.count; ++j)
                layer_step(layers[j], idx, n_axes, wavelength, k2, false, med, where, period);
            total = chain_mul(total, chain_pow(period, seg.repeat - 1));
        }
    }
    // substrate.transfer with thickness 0, _multilayers.py:208-217
    layer_step(layers[P.n_layers - 1], idx, n_axes, wavelength, k2, true, med, where, total);

    // r = M21 / M11, t = 1 / M11, t[~where] = 0   (_multilayers.py:218-222)
    const cplx r_s = total.s.m.c / total.s.m.a, r_p = total.p.m.c / total.p.m.a;
    cplx t_s = total.s.d / (total.s.m.a * total.u), t_p = total.p.d / (total.p.m.a * total.u);
    if (!where) {
        t_s = C(0.0);
        t_p = C(0.0);
    }

    // multilayer_efficiency, _multilayers.py:501-532: the substrate direction from the
    // AMBIENT direction by snells_law_scalar is the invariant's value in the substrate,
    // i.e. what the chain just computed for its last medium
    if (P.r_s) P.r_s[e] = norm2(r_s);
    if (P.r_p) P.r_p[e] = norm2(r_p);
    if (P.t_s) P.t_s[e] = norm2(t_s) * (med.q_s / q_amb_s).re;
    if (P.t_p) P.t_p[e] = norm2(t_p) * (med.q_p / q_amb_p).re;
}

// ---------------------------------------------------------------------------
// Explicit (non-periodic) stacks: only the FIRST COLUMN of the total transfer matrix is
// needed (r = M21 / M11, t = 1 / M11), so the product is evaluated right to left as a
// matrix-vector recursion starting in the substrate: v <- W_j (U_j v), 5 complex products
// per layer and polarisation instead of 10.  The layers may be visited in any order because
// n sin(theta) is invariant (see layer_step).  The reference's top-down overflow mask
// (_layers.py:261-277) is reproduced exactly: at a layer whose |exp(-i beta)| >= 1e10 the
// propagation is dropped and everything BELOW becomes the identity, i.e. the vector is
// reset to (1, 0) before that layer's refraction; the topmost such layer wins and t = 0.
// The scalars only enter |t|^2 and are tracked as real products.
// ---------------------------------------------------------------------------
__device__ __forceinline__ Medium medium_of(const LayerDev& L, const unsigned* idx, int n_axes, cplx k2) {
    long long on = 0;
    if (L.n_axis >= 0) {
        const int a = L.n_axis;
        on = (long long)(a == 0 ? idx[0] : (a == 1 ? idx[1] : (a == 2 ? idx[2] : idx[3]))) *
             (a == 0 ? L.n_stride[0] : (a == 1 ? L.n_stride[1] : (a == 2 ? L.n_stride[2] : L.n_stride[3])));
    } else if (L.n_axis == -2) {
        for (int a = 0; a < n_axes; ++a) on += (long long)idx[a] * L.n_stride[a];
    }
    Medium m;
    m.n = C(__ldg(L.n_re + on), L.n_im ? __ldg(L.n_im + on) : 0.0);
    const cplx inv_n = cinv(m.n);
    m.dir = csqrt(C(1.0) - k2 * (inv_n * inv_n));
    m.q_s = m.dir * m.n;
    m.q_p = conj(m.dir) * inv_n;
    return m;
}

__device__ __forceinline__ long long layer_offset(int axis, const long long* stride, const unsigned* idx, int n_axes) {
    if (axis == -1) return 0;
    if (axis >= 0) {
        const unsigned i = axis == 0 ? idx[0] : (axis == 1 ? idx[1] : (axis == 2 ? idx[2] : idx[3]));
        const long long st = axis == 0 ? stride[0] : (axis == 1 ? stride[1] : (axis == 2 ? stride[2] : stride[3]));
        return (long long)i * st;
    }
    long long o = 0;
    for (int a = 0; a < n_axes; ++a) o += (long long)idx[a] * stride[a];
    return o;
}

struct Column {
    cplx v0, v1;  // first column of the (unnormalised) product so far
    double d2;    // prod |2 q_i|^2 of the refractions applied
};

template <int MINB>
__global__ void __launch_bounds__(128, MINB) multilayer_explicit_kernel(const __grid_constant__ MultilayerParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LayerDev* layers = reinterpret_cast<LayerDev*>(smem_raw);
    {
        const int n_words = P.n_layers * (int)(sizeof(LayerDev) / sizeof(long long));
        const long long* src = reinterpret_cast<const long long*>(P.layers);
        long long* dst = reinterpret_cast<long long*>(layers);
        for (int k = threadIdx.x; k < n_words; k += blockDim.x) dst[k] = src[k];
    }
    __syncthreads();

    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= P.n_eval) return;

    unsigned idx[OPTK_ML_MAX_AXES] = {0, 0, 0, 0};
    {
        uint32_t rem = (uint32_t)e;
        for (int a = P.in.n_axes - 1; a >= 0; --a) {
            if (a == 0) {
                idx[a] = rem;
            } else {
                uint32_t q, r;
                divmod(rem, P.div[a], q, r);
                idx[a] = r;
                rem = q;
            }
        }
    }
    const int n_axes = P.in.n_axes;
    long long ow = 0, od = 0, on = 0;
    for (int a = 0; a < n_axes; ++a) {
        ow += (long long)idx[a] * P.in.wavelength_stride[a];
        od += (long long)idx[a] * P.in.direction_stride[a];
        on += (long long)idx[a] * P.in.n_stride[a];
    }
    const double wavelength = __ldg(P.in.wavelength + ow);
    const cplx dir0 = C(__ldg(P.in.direction_re + od), P.in.direction_im ? __ldg(P.in.direction_im + od) : 0.0);
    const cplx n0 = C(__ldg(P.in.n_re + on), P.in.n_im ? __ldg(P.in.n_im + on) : 0.0);
    Medium ambient;
    ambient.n = n0;
    ambient.dir = dir0;
    ambient.q_s = dir0 * n0;
    ambient.q_p = conj(dir0) / n0;
    const cplx k2 = (n0 * n0) * (C(1.0) - dir0 * dir0);  // (n sin(theta))^2, invariant through the stack
    const double two_pi_over_w = 2.0 * 3.141592653589793 / wavelength;

    // start in the substrate (thickness forced to 0, _multilayers.py:190-193)
    int j_lo = P.n_layers - 1;
    Medium lo = medium_of(layers[j_lo], idx, n_axes, k2);
    const cplx q_sub_s = lo.q_s, q_sub_p = lo.q_p;
    Column cs{C(1.0), C(0.0), 1.0}, cp{C(1.0), C(0.0), 1.0};
    double growth_total = 1.0;  // prod |exp(-i beta_j)|
    bool tripped = false;
    bool lo_is_substrate = true;

    // visit the interfaces from the bottom up; `count` interfaces = number of layers incl. substrate
    for (int g = P.n_segments; g >= 0; --g) {
        // segment g - 1 supplies the upper media; g == 0 is the ambient medium
        const int first = g > 0 ? P.segments[g - 1].first : 0;
        const int count = g > 0 ? P.segments[g - 1].count : 1;
        for (int jj = count - 1; jj >= 0; --jj) {
            const bool up_is_ambient = g == 0;
            const int j_up = first + jj;
            const Medium up = up_is_ambient ? ambient : medium_of(layers[j_up], idx, n_axes, k2);
            const LayerDev& L = layers[j_lo];

            // propagation through the lower layer (matrices.py:241-246)
            if (!lo_is_substrate && L.thickness) {
                const double h = __ldg(L.thickness + layer_offset(L.t_axis, L.t_stride, idx, n_axes));
                const double br = two_pi_over_w * h * lo.q_s.re, bi = two_pi_over_w * h * lo.q_s.im;
                const double growth = fexp(bi);
                const bool ok = growth < 1e10;  // _layers.py:271 (NaN -> false)
                double sn, cn;
                fsincos(br, &sn, &cn);
                const double decay = frcp(growth * growth);
                const cplx w = C(decay * (cn * cn - sn * sn), decay * (2.0 * sn * cn));  // exp(2 i beta)
                // ok: v1 *= w.  not ok: everything below is the identity -> v = (1, 0), scalars reset
                cs.v1 = ok ? cs.v1 * w : C(0.0);
                cp.v1 = ok ? cp.v1 * w : C(0.0);
                cs.v0 = ok ? cs.v0 : C(1.0);
                cp.v0 = ok ? cp.v0 : C(1.0);
                cs.d2 = ok ? cs.d2 : 1.0;
                cp.d2 = ok ? cp.d2 : 1.0;
                growth_total = ok ? growth_total * growth : 1.0;
                tripped = tripped || !ok;
            }

            // refraction at the top interface of the lower layer (matrices.py:127-161); the
            // interface profile belongs to the lower layer and sees the UPPER medium
            double rough_s = 1.0, rough_p = 1.0;
            if (L.profile_kind) {
                const double width = L.width ? __ldg(L.width + layer_offset(L.w_axis, L.w_stride, idx, n_axes)) : 0.0;
                const double k = 2.0 * two_pi_over_w;
                rough_s = interface_factor(L.profile_kind, width, k * up.q_s.re);
                rough_p = interface_factor(L.profile_kind, width, k * (up.n * conj(up.dir)).re);
            }
            {
                const cplx a = up.q_s + lo.q_s, b = rough_s * (up.q_s - lo.q_s);
                const cplx v0 = a * cs.v0 + b * cs.v1, v1 = b * cs.v0 + a * cs.v1;
                cs.v0 = v0;
                cs.v1 = v1;
                cs.d2 *= 4.0 * norm2(up.q_s);
            }
            {
                const cplx a = up.q_p + lo.q_p, b = rough_p * (up.q_p - lo.q_p);
                const cplx v0 = a * cp.v0 + b * cp.v1, v1 = b * cp.v0 + a * cp.v1;
                cp.v0 = v0;
                cp.v1 = v1;
                cp.d2 *= 4.0 * norm2(up.q_p);
            }
            lo = up;
            j_lo = j_up;
            lo_is_substrate = false;
        }
    }

    // R = |M21 / M11|^2, T = |1 / M11|^2 Re(q_substrate / q_ambient)  (_multilayers.py:218-222, 501-532)
    const double m_s = norm2(cs.v0), m_p = norm2(cp.v0);
    const double g2 = growth_total * growth_total;
    if (P.r_s) P.r_s[e] = fdiv(norm2(cs.v1), m_s);
    if (P.r_p) P.r_p[e] = fdiv(norm2(cp.v1), m_p);
    if (P.t_s) P.t_s[e] = tripped ? 0.0 * (q_sub_s / ambient.q_s).re : fdiv(cs.d2, m_s * g2) * (q_sub_s / ambient.q_s).re;
    if (P.t_p) P.t_p[e] = tripped ? 0.0 * (q_sub_p / ambient.q_p).re : fdiv(cp.d2, m_p * g2) * (q_sub_p / ambient.q_p).re;
}

// ---------------------------------------------------------------------------
// Explicit stacks with a configuration axis that only the thicknesses depend on (BASELINE
// cfg 4: the period scaled per configuration): the media, Fresnel terms and interface
// factors depend on (wavelength, angle) only, so one thread carries NC configurations through
// the stack and computes them once per layer; per configuration only the propagation phase
// (one exp, one sincos) and the 2 x 5 complex products remain.
// ---------------------------------------------------------------------------
template <int NC, int MINB>
__global__ void __launch_bounds__(128, MINB) multilayer_reuse_kernel(const __grid_constant__ MultilayerParams P,
                                                                    long long n_outer, int n_groups) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LayerDev* layers = reinterpret_cast<LayerDev*>(smem_raw);
    {
        const int n_words = P.n_layers * (int)(sizeof(LayerDev) / sizeof(long long));
        const long long* src = reinterpret_cast<const long long*>(P.layers);
        long long* dst = reinterpret_cast<long long*>(layers);
        for (int k = threadIdx.x; k < n_words; k += blockDim.x) dst[k] = src[k];
    }
    __syncthreads();

    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n_outer * n_groups) return;
    const int ra = P.reuse_axis;
    const int n_axes = P.in.n_axes;
    const long long eo = tid % n_outer;  // index over the other axes (adjacent threads: adjacent outer points)
    const int c0 = (int)(tid / n_outer) * NC;
    const int n_c = (int)P.in.dims[ra];

    // decompose eo over the axes other than `ra` (C order), idx[ra] = 0; dense output offset
    unsigned idx[OPTK_ML_MAX_AXES] = {0, 0, 0, 0};
    long long out_base = 0, out_stride_c = 1;
    {
        long long rem = eo, dense = 1;
        for (int a = n_axes - 1; a >= 0; --a) {
            if (a == ra) {
                out_stride_c = dense;
            } else {
                const long long d = P.in.dims[a];
                idx[a] = (unsigned)(rem % d);
                rem /= d;
                out_base += idx[a] * dense;
            }
            dense *= P.in.dims[a];
        }
    }
    long long ow = 0, od = 0, on = 0;
    for (int a = 0; a < n_axes; ++a) {
        ow += (long long)idx[a] * P.in.wavelength_stride[a];
        od += (long long)idx[a] * P.in.direction_stride[a];
        on += (long long)idx[a] * P.in.n_stride[a];
    }
    const double wavelength = __ldg(P.in.wavelength + ow);
    const cplx dir0 = C(__ldg(P.in.direction_re + od), P.in.direction_im ? __ldg(P.in.direction_im + od) : 0.0);
    const cplx n0 = C(__ldg(P.in.n_re + on), P.in.n_im ? __ldg(P.in.n_im + on) : 0.0);
    Medium ambient;
    ambient.n = n0;
    ambient.dir = dir0;
    ambient.q_s = dir0 * n0;
    ambient.q_p = conj(dir0) / n0;
    const cplx k2 = (n0 * n0) * (C(1.0) - dir0 * dir0);
    const double two_pi_over_w = 2.0 * 3.141592653589793 / wavelength;

    int j_lo = P.n_layers - 1;
    Medium lo = medium_of(layers[j_lo], idx, n_axes, k2);
    const cplx q_sub_s = lo.q_s, q_sub_p = lo.q_p;
    Column cs[NC], cp[NC];
    double growth_total[NC];
    bool tripped[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        cs[c] = Column{C(1.0), C(0.0), 1.0};
        cp[c] = Column{C(1.0), C(0.0), 1.0};
        growth_total[c] = 1.0;
        tripped[c] = false;
    }
    bool lo_is_substrate = true;

    for (int g = P.n_segments; g >= 0; --g) {
        const int first = g > 0 ? P.segments[g - 1].first : 0;
        const int count = g > 0 ? P.segments[g - 1].count : 1;
        for (int jj = count - 1; jj >= 0; --jj) {
            const bool up_is_ambient = g == 0;
            const int j_up = first + jj;
            const Medium up = up_is_ambient ? ambient : medium_of(layers[j_up], idx, n_axes, k2);
            const LayerDev& L = layers[j_lo];

            // per configuration: propagation through the lower layer
            if (!lo_is_substrate && L.thickness) {
                const long long t0 = layer_offset(L.t_axis == ra ? -1 : L.t_axis, L.t_stride, idx, n_axes);
                const long long tc = L.t_stride[ra];
                const double kr = two_pi_over_w * lo.q_s.re, ki = two_pi_over_w * lo.q_s.im;
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const int cc = min(c0 + c, n_c - 1);  // lanes past the end repeat the last configuration
                    const double h = __ldg(L.thickness + t0 + (long long)cc * tc);
                    const double growth = fexp(ki * h);
                    const bool ok = growth < 1e10;
                    double sn, cn;
                    fsincos(kr * h, &sn, &cn);
                    const double decay = frcp(growth * growth);
                    const cplx w = C(decay * (cn * cn - sn * sn), decay * (2.0 * sn * cn));
                    cs[c].v1 = ok ? cs[c].v1 * w : C(0.0);
                    cp[c].v1 = ok ? cp[c].v1 * w : C(0.0);
                    cs[c].v0 = ok ? cs[c].v0 : C(1.0);
                    cp[c].v0 = ok ? cp[c].v0 : C(1.0);
                    cs[c].d2 = ok ? cs[c].d2 : 1.0;
                    cp[c].d2 = ok ? cp[c].d2 : 1.0;
                    growth_total[c] = ok ? growth_total[c] * growth : 1.0;
                    tripped[c] = tripped[c] || !ok;
                }
            }

            // shared by all configurations: refraction at the top interface of the lower layer
            double rough_s = 1.0, rough_p = 1.0;
            if (L.profile_kind) {
                const double width = L.width ? __ldg(L.width + layer_offset(L.w_axis, L.w_stride, idx, n_axes)) : 0.0;
                const double k = 2.0 * two_pi_over_w;
                rough_s = interface_factor(L.profile_kind, width, k * up.q_s.re);
                rough_p = interface_factor(L.profile_kind, width, k * (up.n * conj(up.dir)).re);
            }
            const cplx a_s = up.q_s + lo.q_s, b_s = rough_s * (up.q_s - lo.q_s);
            const cplx a_p = up.q_p + lo.q_p, b_p = rough_p * (up.q_p - lo.q_p);
            const double f_s = 4.0 * norm2(up.q_s), f_p = 4.0 * norm2(up.q_p);
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const cplx s0 = a_s * cs[c].v0 + b_s * cs[c].v1, s1 = b_s * cs[c].v0 + a_s * cs[c].v1;
                cs[c].v0 = s0;
                cs[c].v1 = s1;
                cs[c].d2 *= f_s;
                const cplx p0 = a_p * cp[c].v0 + b_p * cp[c].v1, p1 = b_p * cp[c].v0 + a_p * cp[c].v1;
                cp[c].v0 = p0;
                cp[c].v1 = p1;
                cp[c].d2 *= f_p;
            }
            lo = up;
            j_lo = j_up;
            lo_is_substrate = false;
        }
    }

    const double ratio_s = (q_sub_s / ambient.q_s).re, ratio_p = (q_sub_p / ambient.q_p).re;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        if (c0 + c < n_c) {
            const long long e = out_base + (long long)(c0 + c) * out_stride_c;
            const double m_s = norm2(cs[c].v0), m_p = norm2(cp[c].v0);
            const double g2 = growth_total[c] * growth_total[c];
            if (P.r_s) P.r_s[e] = fdiv(norm2(cs[c].v1), m_s);
            if (P.r_p) P.r_p[e] = fdiv(norm2(cp[c].v1), m_p);
            if (P.t_s) P.t_s[e] = tripped[c] ? 0.0 * ratio_s : fdiv(cs[c].d2, m_s * g2) * ratio_s;
            if (P.t_p) P.t_p[e] = tripped[c] ? 0.0 * ratio_p : fdiv(cp[c].d2, m_p * g2) * ratio_p;
        }
    }
}

int launch_multilayer(const MultilayerParams& P, cudaStream_t stream) {
    if (P.n_eval <= 0) return OPTK_OK;
    const int block = 128;
    const long long grid = (P.n_eval + block - 1) / block;
    if (grid > 0x7fffffffLL) {
        set_error("optk_multilayer: too many evaluations for one launch (%lld)", P.n_eval);
        return OPTK_ERR_INVALID;
    }
    const size_t smem = (size_t)P.n_layers * sizeof(LayerDev);
    bool periodic = false;
    for (int g = 0; g < P.n_segments; ++g) periodic = periodic || P.segments[g].repeat > 1;
    static const int occ = [] {
        const char* e = getenv("OPTK_ML_OCC");
        return e ? atoi(e) : 4;
    }();
    static const int reuse = [] {
        const char* e = getenv("OPTK_ML_REUSE");
        return e ? atoi(e) : 4;
    }();
    if (!periodic && P.reuse_axis >= 0 && reuse > 1) {
        const int nc = reuse == 42 ? 4 : (reuse == 32 ? 3 : reuse);
        const long long n_c = P.in.dims[P.reuse_axis];
        const long long n_outer = P.n_eval / n_c;
        const int n_groups = (int)((n_c + nc - 1) / nc);
        const long long threads = n_outer * n_groups;
        const long long g2 = (threads + block - 1) / block;
        if (reuse == 42)
            multilayer_reuse_kernel<4, 2><<<(unsigned)g2, block, smem, stream>>>(P, n_outer, n_groups);
        else if (reuse == 32)
            multilayer_reuse_kernel<3, 2><<<(unsigned)g2, block, smem, stream>>>(P, n_outer, n_groups);
        else if (reuse == 3)
            multilayer_reuse_kernel<3, 3><<<(unsigned)g2, block, smem, stream>>>(P, n_outer, n_groups);
        else if (reuse == 2)
            multilayer_reuse_kernel<2, 4><<<(unsigned)g2, block, smem, stream>>>(P, n_outer, n_groups);
        else
            multilayer_reuse_kernel<4, 3><<<(unsigned)g2, block, smem, stream>>>(P, n_outer, n_groups);
        OPTK_CUDA(cudaGetLastError());
        return OPTK_OK;
    }
    if (periodic)
        multilayer_kernel<true, 2><<<(unsigned)grid, block, smem, stream>>>(P);
    else if (occ == 0)
        multilayer_kernel<false, 4><<<(unsigned)grid, block, smem, stream>>>(P);  // matrix form, for A/B tests
    else if (occ == 3)
        multilayer_explicit_kernel<3><<<(unsigned)grid, block, smem, stream>>>(P);
    else if (occ == 5)
        multilayer_explicit_kernel<5><<<(unsigned)grid, block, smem, stream>>>(P);
    else
        multilayer_explicit_kernel<4><<<(unsigned)grid, block, smem, stream>>>(P);
    OPTK_CUDA(cudaGetLastError());
    return OPTK_OK;
}

}  // namespace optk
