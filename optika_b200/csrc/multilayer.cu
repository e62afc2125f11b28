// Kernel 3: transfer-matrix (Yeh) efficiency of a multilayer stack (sm_100a, fp64).
//
// One thread evaluates one point of the (wavelength x angle x configuration)
// grid.  The layer table is staged in shared memory; the complex 2x2 chain for
// both polarisations stays in registers.  Restates
// optika/materials/_multilayers.py:187-237, 485-532, _layers.py:243-277, 487-499,
// 625-645, matrices.py:127-161, 241-246 and profiles.py:103-126.
#include "common.cuh"
#include "params.cuh"

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
    const double s = 1.0 / norm2(a);
    return C(a.re * s, -a.im * s);
}
__device__ __forceinline__ cplx operator/(cplx a, cplx b) { return a * cinv(b); }

// principal square root with the C99 / numpy branch conventions
__device__ __forceinline__ cplx csqrt(cplx z) {
    if (z.re == 0.0 && z.im == 0.0) return C(0.0, z.im);
    const double m = sqrt(norm2(z));
    if (z.re >= 0.0) {
        const double t = sqrt(0.5 * (m + z.re));
        return C(t, z.im / (2.0 * t));
    }
    const double t = sqrt(0.5 * (m - z.re));
    return C(fabs(z.im) / (2.0 * t), copysign(t, z.im));
}

struct Mat2 {
    cplx a, b, c, d;  // [[a, b], [c, d]]
};

__device__ __forceinline__ Mat2 identity2() { return Mat2{C(1.0), C(0.0), C(0.0), C(1.0)}; }

__device__ __forceinline__ Mat2 matmul(const Mat2& x, const Mat2& y) {
    return Mat2{x.a * y.a + x.b * y.c, x.a * y.b + x.b * y.d, x.c * y.a + x.d * y.c, x.c * y.b + x.d * y.d};
}

// n-fold product by square-and-multiply (Cartesian2dMatrixArray.power, _layers.py:643)
__device__ __forceinline__ Mat2 matpow(Mat2 x, int n) {
    Mat2 r = identity2();
    while (n > 0) {
        if (n & 1) r = matmul(r, x);
        n >>= 1;
        if (n) x = matmul(x, x);
    }
    return r;
}

// profiles.py:103-126 + the four _derivative_fourier_transform bodies
__device__ __forceinline__ double interface_factor(int kind, double width, double s) {
    const double pi = 3.141592653589793;
    switch (kind) {
        case 1: {
            const double x = s * width;
            return exp(-(x * x) / 2.0);
        }
        case 2: {
            const double x = s * width;
            return 1.0 / (1.0 + (x * x) / 2.0);
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

struct Chain {
    cplx n, dir;      // medium the light is currently in, cosine of its propagation angle
    cplx q_s, q_p;    // n cos(theta),  conj(cos(theta)) / n
    bool where;
};

// Layer.transfer (_layers.py:229-277) for both polarisations; `is_substrate`
// forces the thickness to zero (_multilayers.py:190-193).
__device__ __forceinline__ void layer_transfer(const LayerDev& L, const unsigned* idx, int n_axes, double wavelength,
                                               bool is_substrate, Chain& st, Mat2& t_s, Mat2& t_p) {
    long long on = 0, ot = 0, ow = 0;
    for (int a = 0; a < n_axes; ++a) {
        on += (long long)idx[a] * L.n_stride[a];
        ot += (long long)idx[a] * L.t_stride[a];
        ow += (long long)idx[a] * L.w_stride[a];
    }
    const cplx n_j = C(__ldg(L.n_re + on), L.n_im ? __ldg(L.n_im + on) : 0.0);
    const double h = (is_substrate || !L.thickness) ? 0.0 : __ldg(L.thickness + ot);

    // snells_law_scalar, _snells_law.py:34-38
    const cplx sin_i = csqrt(C(1.0) - st.dir * st.dir);
    const cplx sin_t = st.n * sin_i / n_j;
    const cplx dir_j = csqrt(C(1.0) - sin_t * sin_t);

    // matrices.refraction, matrices.py:127-161
    const cplx q_s_j = dir_j * n_j;
    const cplx q_p_j = conj(dir_j) / n_j;
    double rough_s = 1.0, rough_p = 1.0;
    if (L.profile_kind) {
        // s = Re(4 pi n_i direction_i / wavelength); direction_i is conjugated for p
        const double width = L.width ? __ldg(L.width + ow) : 0.0;
        const double k = 4.0 * 3.141592653589793 / wavelength;
        rough_s = interface_factor(L.profile_kind, width, k * (st.n * st.dir).re);
        rough_p = interface_factor(L.profile_kind, width, k * (st.n * conj(st.dir)).re);
    }
    Mat2 w_s, w_p;
    {
        const cplx a = st.q_s + q_s_j;
        const cplx r = rough_s * ((st.q_s - q_s_j) / a);
        const cplx it = cinv((2.0 * st.q_s) / a);
        w_s = Mat2{it, r * it, r * it, it};
    }
    {
        const cplx a = st.q_p + q_p_j;
        const cplx r = rough_p * ((st.q_p - q_p_j) / a);
        const cplx it = cinv((2.0 * st.q_p) / a);
        w_p = Mat2{it, r * it, r * it, it};
    }
    if (!st.where) {  // refraction = where(where, refraction, identity), _layers.py:261-262
        w_s = identity2();
        w_p = identity2();
    }

    // matrices.propagation, matrices.py:241-246: beta = 2 pi h n cos(theta) / wavelength
    const cplx beta = (2.0 * 3.141592653589793 * h / wavelength) * (n_j * dir_j);
    double sn, cs;
    sincos(beta.re, &sn, &cs);
    const double ep = exp(beta.im), em = exp(-beta.im);
    const cplx u00 = C(ep * cs, -ep * sn);  // exp(-i beta)
    const cplx u11 = C(em * cs, em * sn);   // exp(+i beta)
    const bool where_propagation = sqrt(norm2(u00)) < 1e10;  // _layers.py:271
    st.where = st.where && where_propagation;
    if (st.where) {
        t_s = Mat2{w_s.a * u00, w_s.b * u11, w_s.c * u00, w_s.d * u11};
        t_p = Mat2{w_p.a * u00, w_p.b * u11, w_p.c * u00, w_p.d * u11};
    } else {
        t_s = w_s;
        t_p = w_p;
    }
    st.n = n_j;
    st.dir = dir_j;
    st.q_s = q_s_j;
    st.q_p = q_p_j;
}

__global__ void __launch_bounds__(128) multilayer_kernel(const __grid_constant__ MultilayerParams P) {
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

    Chain st;
    st.n = n0;
    st.dir = dir0;
    st.q_s = dir0 * n0;
    st.q_p = conj(dir0) / n0;
    st.where = true;
    const cplx q_amb_s = st.q_s, q_amb_p = st.q_p;

    Mat2 m_s = identity2(), m_p = identity2();
    const int n_axes = P.in.n_axes;

    for (int g = 0; g < P.n_segments; ++g) {
        const optk_ml_segment_t seg = P.segments[g];
        if (seg.repeat == 1) {
            // LayerSequence.transfer, _layers.py:487-499
            for (int j = seg.first; j < seg.first + seg.count; ++j) {
                Mat2 t_s, t_p;
                layer_transfer(layers[j], idx, n_axes, wavelength, false, st, t_s, t_p);
                m_s = matmul(m_s, t_s);
                m_p = matmul(m_p, t_p);
            }
        } else if (seg.repeat > 1) {
            // PeriodicLayerSequence.transfer, _layers.py:625-645: the first period
            // explicitly, the second period raised to the power (repeat - 1)
            Mat2 start_s = identity2(), start_p = identity2();
            for (int j = seg.first; j < seg.first + seg.count; ++j) {
                Mat2 t_s, t_p;
                layer_transfer(layers[j], idx, n_axes, wavelength, false, st, t_s, t_p);
                start_s = matmul(start_s, t_s);
                start_p = matmul(start_p, t_p);
            }
            Mat2 per_s = identity2(), per_p = identity2();
            for (int j = seg.first; j < seg.first + seg.count; ++j) {
                Mat2 t_s, t_p;
                layer_transfer(layers[j], idx, n_axes, wavelength, false, st, t_s, t_p);
                per_s = matmul(per_s, t_s);
                per_p = matmul(per_p, t_p);
            }
            m_s = matmul(m_s, matmul(start_s, matpow(per_s, seg.repeat - 1)));
            m_p = matmul(m_p, matmul(start_p, matpow(per_p, seg.repeat - 1)));
        }
    }
    // substrate.transfer with thickness 0, _multilayers.py:208-217
    const LayerDev& sub = layers[P.n_layers - 1];
    {
        Mat2 t_s, t_p;
        layer_transfer(sub, idx, n_axes, wavelength, true, st, t_s, t_p);
        m_s = matmul(m_s, t_s);
        m_p = matmul(m_p, t_p);
    }

    // r = M21 / M11, t = 1 / M11, t[~where] = 0   (_multilayers.py:218-222)
    const cplx r_s = m_s.c / m_s.a, r_p = m_p.c / m_p.a;
    cplx t_s = cinv(m_s.a), t_p = cinv(m_p.a);
    if (!st.where) {
        t_s = C(0.0);
        t_p = C(0.0);
    }

    // multilayer_efficiency, _multilayers.py:501-532: the substrate direction is
    // recomputed from the AMBIENT direction by snells_law_scalar
    const cplx n_sub = st.n;
    const cplx sin_i = csqrt(C(1.0) - dir0 * dir0);
    const cplx sin_t = n0 * sin_i / n_sub;
    const cplx dir_sub = csqrt(C(1.0) - sin_t * sin_t);
    const cplx q_sub_s = dir_sub * n_sub;
    const cplx q_sub_p = conj(dir_sub) / n_sub;

    if (P.r_s) P.r_s[e] = norm2(r_s);
    if (P.r_p) P.r_p[e] = norm2(r_p);
    if (P.t_s) P.t_s[e] = norm2(t_s) * (q_sub_s / q_amb_s).re;
    if (P.t_p) P.t_p[e] = norm2(t_p) * (q_sub_p / q_amb_p).re;
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
    multilayer_kernel<<<(unsigned)grid, block, smem, stream>>>(P);
    OPTK_CUDA(cudaGetLastError());
    return OPTK_OK;
}

}  // namespace optk
