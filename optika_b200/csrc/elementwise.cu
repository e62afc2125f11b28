// Small per-ray kernels of the multilayer-coated-surface chain (include/optk.h, "per-ray
// multilayer efficiency"): table interpolation of optical constants at the ray wavelengths
// and the polarisation-averaged efficiency applied to the intensity.  HBM-bound streams.
#include "common.cuh"
#include "bin.cuh"
#include "params.cuh"

namespace optk {

// numpy.interp: linear, ends clamped, NaN -> NaN (Chemical.n, optika/chemicals/_chemicals.py:136-142)
__global__ void __launch_bounds__(256)
interp_kernel(long long n, const double* __restrict__ x, int m, const double* __restrict__ xp,
              const double* __restrict__ fp_re, const double* __restrict__ fp_im, double* __restrict__ out_re,
              double* __restrict__ out_im) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = __ldg(x + i);
    double re, im = 0.0;
    if (v != v) {
        re = im = v;
    } else if (m == 1 || v <= __ldg(xp)) {
        re = __ldg(fp_re);
        if (fp_im) im = __ldg(fp_im);
    } else if (v >= __ldg(xp + m - 1)) {
        re = __ldg(fp_re + m - 1);
        if (fp_im) im = __ldg(fp_im + m - 1);
    } else {
        int lo = 0, hi = m - 1;  // xp[lo] <= v < xp[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (v >= __ldg(xp + mid)) lo = mid; else hi = mid;
        }
        const double x0 = __ldg(xp + lo), dx = __ldg(xp + lo + 1) - x0;
        const double r0 = __ldg(fp_re + lo);
        re = (__ldg(fp_re + lo + 1) - r0) / dx * (v - x0) + r0;
        if (fp_im) {
            const double i0 = __ldg(fp_im + lo);
            im = (__ldg(fp_im + lo + 1) - i0) / dx * (v - x0) + i0;
        }
    }
    out_re[i] = re;
    if (out_im) out_im[i] = im;
}

__global__ void __launch_bounds__(256)
apply_efficiency_kernel(long long n, double* __restrict__ intensity, const double* __restrict__ e_s,
                        const double* __restrict__ e_p) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) intensity[i] *= (e_s[i] + e_p[i]) / 2;
}

static int grid_for(long long n, unsigned* grid) {
    const long long g = (n + 255) / 256;
    if (g > 0x7fffffffLL) {
        set_error("too many elements for one launch (%lld)", n);
        return OPTK_ERR_INVALID;
    }
    *grid = (unsigned)g;
    return OPTK_OK;
}

int launch_interp(long long n, const double* x, int m, const double* xp, const double* fp_re, const double* fp_im,
                  double* out_re, double* out_im, cudaStream_t stream) {
    if (n == 0) return OPTK_OK;
    unsigned grid;
    int rc = grid_for(n, &grid);
    if (rc) return rc;
    interp_kernel<<<grid, 256, 0, stream>>>(n, x, m, xp, fp_re, fp_im, out_re, out_im);
    OPTK_CUDA(cudaGetLastError());
    return OPTK_OK;
}

int launch_apply_efficiency(long long n, double* intensity, const double* e_s, const double* e_p, cudaStream_t stream) {
    if (n == 0) return OPTK_OK;
    unsigned grid;
    int rc = grid_for(n, &grid);
    if (rc) return rc;
    apply_efficiency_kernel<<<grid, 256, 0, stream>>>(n, intensity, e_s, e_p);
    OPTK_CUDA(cudaGetLastError());
    return OPTK_OK;
}

// optk_debug_math: the fp64 helper sequences of common.cuh on arrays
__global__ void __launch_bounds__(256)
math_kernel(int op, long long n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = a[i];
    double y;
    switch (op) {
        case 0: y = fdiv(x, b[i]); break;
        case 1: y = frcp(x); break;
        case 2: y = fsqrt(x); break;
        case 3: y = frsqrt(x); break;
        case 4: y = frcp_raw(x); break;
        case 6: y = fdiv_finite(x, b[i]); break;
        case 7: y = fdiv_newton(x, b[i]); break;
        default: y = frsqrt_raw(x); break;
    }
    out[i] = y;
}

int launch_debug_math(int op, long long n, const double* a, const double* b, double* out, cudaStream_t stream) {
    if (n == 0) return OPTK_OK;
    unsigned grid;
    int rc = grid_for(n, &grid);
    if (rc) return rc;
    math_kernel<<<grid, 256, 0, stream>>>(op, n, a, b, out);
    OPTK_CUDA(cudaGetLastError());
    return OPTK_OK;
}

// ---------------------------------------------------------------------------
// Reductions over the pupil of traced rays, per field point: what
// SequentialSystem.distortion / vignetting / area_effective take from the ray arrays
// (optika/systems/_sequential.py:1266-1285, 1351-1368, 1501-1506) -- `unvignetted.any(axis_pupil)`,
// `mean(position.xy, axis_pupil, where=unvignetted | ~where)`, `unvignetted.mean(axis_pupil)`,
// `intensity.sum(axis_pupil, where=unvignetted)` -- computed where the rays are.
// Rays are dense arrays with the pupil axes innermost: group g = rays [g n_inner, (g + 1) n_inner).
// One thread per ray; the warp network of bin.cuh merges the rays of a warp by group before the
// reductions go to L2; CTAs visit the rays in the strided order of the fused image kernels.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
reduce_groups_kernel(long long n, FastDiv div_inner, const double* __restrict__ x, const double* __restrict__ y,
                     const double* __restrict__ intensity, const uint8_t* __restrict__ unvignetted,
                     double* sum_intensity, double* sum_x, double* sum_y, unsigned long long* count, double* sum_x_all,
                     double* sum_y_all, int cta_rows, int cta_count) {
    unsigned block = blockIdx.x;
    if (cta_rows) {
        block = (block & 511u) * (unsigned)cta_rows + (block >> 9);
        if (block >= (unsigned)cta_count) return;
    }
    const long long i = (long long)block * blockDim.x + threadIdx.x;
    const bool valid = i < n;
    double px = 0.0, py = 0.0, in = 0.0;
    bool unv = false;
    int group = -1;
    if (valid) {
        uint32_t q, r;
        divmod((uint32_t)i, div_inner, q, r);
        group = (int)q;
        px = __ldg(x + i);
        py = __ldg(y + i);
        in = intensity ? __ldg(intensity + i) : 1.0;
        unv = unvignetted ? (__ldg(unvignetted + i) != 0) : true;
    }
    ImageDev kept = {}, all = {};
    kept.flux = sum_intensity;
    kept.moment_real = sum_x;
    kept.moment_imag = sum_y;
    kept.counts = count;
    all.moment_real = sum_x_all;
    all.moment_imag = sum_y_all;
    image_add(kept, unv ? group : -1, in, px, py, 1u);
    if (sum_x_all || sum_y_all) image_add(all, group, 0.0, px, py, 0u);
}

int launch_reduce_groups(long long n_groups, long long n_inner, const double* x, const double* y, const double* intensity,
                         const uint8_t* unvignetted, double* sum_intensity, double* sum_x, double* sum_y,
                         unsigned long long* count, double* sum_x_all, double* sum_y_all, cudaStream_t stream) {
    const long long n = n_groups * n_inner;
    if (n <= 0) return OPTK_OK;
    const int block = 256;
    long long grid = (n + block - 1) / block;
    const long long rows = (grid + 511) / 512;
    int cta_rows = 0;
    const int cta_count = (int)grid;
    if (grid > 512) {
        cta_rows = (int)rows;
        grid = rows * 512;
    }
    reduce_groups_kernel<<<(unsigned)grid, block, 0, stream>>>(n, make_fastdiv((uint32_t)n_inner), x, y, intensity,
                                                               unvignetted, sum_intensity, sum_x, sum_y, count, sum_x_all,
                                                               sum_y_all, cta_rows, cta_count);
    OPTK_CUDA(cudaGetLastError());
    return OPTK_OK;
}

}  // namespace optk
