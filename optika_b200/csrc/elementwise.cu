// Small per-ray kernels of the multilayer-coated-surface chain (include/optk.h, "per-ray
// multilayer efficiency"): table interpolation of optical constants at the ray wavelengths
// and the polarisation-averaged efficiency applied to the intensity.  HBM-bound streams.
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

}  // namespace optk
