// Detector binning shared by the fused trace kernel and the standalone bin kernel.
//
// Restates AbstractImagingSensor.collect (optika/sensors/_sensors.py:139-161):
// three weighted histograms over (wavelength, x, y) with numpy.histogramdd edge
// semantics -- bin i holds e_i <= v < e_{i+1}, the last bin is closed on the
// right, NaN and out-of-range samples are dropped.  The lookup compares against
// the caller's exact edge VALUES (linspace(bound_lower, bound_upper, n + 1));
// the multiply by 1/width is only a first guess that is then corrected.
#pragma once
#include "common.cuh"

namespace optk {

struct ImageDev {
    int32_t n_w, n_x, n_y;
    int32_t pad;
    const double* e_w;
    const double* e_x;
    const double* e_y;
    double* flux;
    double* moment_real;
    double* moment_imag;
    unsigned long long* counts;
};

// first edge and 1 / mean bin width of the pixel axes: the first guess of the lookup
struct ImageGuess {
    double x0, inv_dx, y0, inv_dy;
};

// Computed once per block into shared memory (edges live in device memory).
__device__ __forceinline__ void image_guess_init(const ImageDev& im, ImageGuess* g) {
    if (threadIdx.x == 0) {
        const double x0 = __ldg(im.e_x), x1 = __ldg(im.e_x + im.n_x);
        const double y0 = __ldg(im.e_y), y1 = __ldg(im.e_y + im.n_y);
        g->x0 = x0;
        g->inv_dx = (double)im.n_x / (x1 - x0);
        g->y0 = y0;
        g->inv_dy = (double)im.n_y / (y1 - y0);
    }
    __syncthreads();
}

// Uniform-ish edges: guess then correct against the exact edge values.
__device__ __forceinline__ int find_bin_guess(const double* __restrict__ e, int n, double v, double e0,
                                              double inv_d) {
    if (!(v >= __ldg(e)) || !(v <= __ldg(e + n))) return -1;  // also rejects NaN
    double g = (v - e0) * inv_d;
    int i = (int)g;
    i = i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
    while (i > 0 && v < __ldg(e + i)) --i;
    while (i < n - 1 && v >= __ldg(e + i + 1)) ++i;
    return i;
}

// Arbitrary monotonic edges (wavelength): binary search, searchsorted(side="right") - 1.
__device__ __forceinline__ int find_bin_search(const double* __restrict__ e, int n, double v) {
    if (!(v >= __ldg(e)) || !(v <= __ldg(e + n))) return -1;
    int lo = 0, hi = n;  // invariant: e[lo] <= v, (hi == n or v < e[hi])
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (v >= __ldg(e + mid)) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Must be called by all 32 lanes of a converged warp.  `bin` < 0 means "drop".
// Warp-aggregated: lanes that hit the same bin are summed in registers and one
// lane issues the global reductions (RED.ADD.F64 / RED.ADD.U64 at L2).
__device__ __forceinline__ void image_add(const ImageDev& im, long long bin, double w_flux, double w_real,
                                          double w_imag) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    unsigned peers = __match_any_sync(full, bin);
    if (peers == full) {
        if (bin < 0) return;
        double f = warp_sum(w_flux);
        double r = im.moment_real ? warp_sum(w_real) : 0.0;
        double m = im.moment_imag ? warp_sum(w_imag) : 0.0;
        if (lane == 0) {
            if (im.flux) atomicAdd(im.flux + bin, f);
            if (im.moment_real) atomicAdd(im.moment_real + bin, r);
            if (im.moment_imag) atomicAdd(im.moment_imag + bin, m);
            if (im.counts) atomicAdd(im.counts + bin, 32ull);
        }
        return;
    }
    // mixed bins: the lowest lane of each peer group gathers its group serially
    const int leader = __ffs(peers) - 1;
    const int n_peers = __popc(peers);
    const int max_peers = __reduce_max_sync(full, (unsigned)n_peers);
    if (max_peers == 1) {
        if (bin >= 0) {
            if (im.flux) atomicAdd(im.flux + bin, w_flux);
            if (im.moment_real) atomicAdd(im.moment_real + bin, w_real);
            if (im.moment_imag) atomicAdd(im.moment_imag + bin, w_imag);
            if (im.counts) atomicAdd(im.counts + bin, 1ull);
        }
        return;
    }
    double f = 0.0, r = 0.0, m = 0.0;
    for (int k = 0; k < 32; ++k) {
        double fk = __shfl_sync(full, w_flux, k);
        double rk = __shfl_sync(full, w_real, k);
        double mk = __shfl_sync(full, w_imag, k);
        if ((peers >> k) & 1u) {
            f += fk;
            r += rk;
            m += mk;
        }
    }
    if (lane == leader && bin >= 0) {
        if (im.flux) atomicAdd(im.flux + bin, f);
        if (im.moment_real) atomicAdd(im.moment_real + bin, r);
        if (im.moment_imag) atomicAdd(im.moment_imag + bin, m);
        if (im.counts) atomicAdd(im.counts + bin, (unsigned long long)n_peers);
    }
}

// Bin one ray given in sensor-local coordinates.  All lanes of the warp call this.
__device__ __forceinline__ void image_bin_ray(const ImageDev& im, const ImageGuess& g, bool valid, double wavelength, double x,
                                              double y, double cos_real, double cos_imag, double intensity,
                                              bool unvignetted) {
    long long bin = -1;
    // flux = intensity * where (optika/sensors/_sensors.py:139): vignetted rays are
    // binned with weight zero and are not counted.
    double flux = unvignetted ? intensity : 0.0 * intensity;
    if (valid) {
        int iw = find_bin_search(im.e_w, im.n_w, wavelength);
        int ix = find_bin_guess(im.e_x, im.n_x, x, g.x0, g.inv_dx);
        int iy = find_bin_guess(im.e_y, im.n_y, y, g.y0, g.inv_dy);
        if (iw >= 0 && ix >= 0 && iy >= 0 && unvignetted)
            bin = ((long long)iw * im.n_x + ix) * im.n_y + iy;
    }
    image_add(im, bin, flux, flux * cos_real, flux * cos_imag);
}

}  // namespace optk
