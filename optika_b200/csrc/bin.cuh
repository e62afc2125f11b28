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
    int32_t group;  // optk_image_t::group_size > 0: bins are groups of consecutive rays, see div_group
    const double* e_w;
    const double* e_x;
    const double* e_y;
    double* flux;
    double* moment_real;
    double* moment_imag;
    unsigned long long* counts;
    int32_t has_range;  // range[] below is valid: no loads needed for the guess
    int32_t pad2;
    double range[6];    // first / last edge of wavelength, x, y
    FastDiv div_group;  // divisor group_size
};

// first / last edge and 1 / mean bin width of the pixel axes: range test and first guess
struct ImageGuess {
    double x0, x1, inv_dx, y0, y1, inv_dy, w0, w1;
};

// Computed once per block into shared memory by ONE thread (the caller places the barrier).
// Edges live in device memory; with the caller's range hint no load is needed.
__device__ __forceinline__ void image_guess_fill(const ImageDev& im, ImageGuess* g) {
    double w0, w1, x0, x1, y0, y1;
    if (im.has_range) {
        w0 = im.range[0]; w1 = im.range[1];
        x0 = im.range[2]; x1 = im.range[3];
        y0 = im.range[4]; y1 = im.range[5];
    } else {
        w0 = __ldg(im.e_w); w1 = __ldg(im.e_w + im.n_w);
        x0 = __ldg(im.e_x); x1 = __ldg(im.e_x + im.n_x);
        y0 = __ldg(im.e_y); y1 = __ldg(im.e_y + im.n_y);
    }
    g->x0 = x0;
    g->x1 = x1;
    g->inv_dx = (double)im.n_x / (x1 - x0);
    g->y0 = y0;
    g->y1 = y1;
    g->inv_dy = (double)im.n_y / (y1 - y0);
    g->w0 = w0;
    g->w1 = w1;
}

__device__ __forceinline__ void image_guess_init(const ImageDev& im, ImageGuess* g) {
    if (threadIdx.x == 0) image_guess_fill(im, g);
    __syncthreads();
}

// Uniform-ish edges: guess then correct against the exact edge values.
// numpy.histogramdd: bin i holds e[i] <= v < e[i+1]; the last bin also holds v == e[n].
__device__ __forceinline__ int find_bin_guess(const double* __restrict__ e, int n, double v, double e0, double e1,
                                              double inv_d) {
    if (!(v >= e0) || !(v <= e1)) return -1;  // also rejects NaN
    int i = (int)((v - e0) * inv_d);
    i = i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
    double lo = __ldg(e + i), hi = __ldg(e + i + 1);
    // the guess is off by at most a bin for linspace edges; loop only in the rare other cases
    while (v < lo) {
        --i;
        hi = lo;
        lo = __ldg(e + i);
    }
    while (v >= hi && i < n - 1) {
        ++i;
        lo = hi;
        hi = __ldg(e + i + 1);
    }
    return i;
}

// Arbitrary monotonic edges (wavelength): binary search, searchsorted(side="right") - 1.
__device__ __forceinline__ int find_bin_search(const double* __restrict__ e, int n, double v, double e0, double e1) {
    if (!(v >= e0) || !(v <= e1)) return -1;
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
// Warp aggregation by a butterfly "merge if equal" network: at each of the five steps a
// lane and its partner (lane ^ o) compare bins; when they agree the lower lane absorbs the
// partner's sums and the partner retires.  The rays of a warp are neighbours in the pupil, so
// they land on a few adjacent pixels in runs and most lanes retire; whoever is left issues
// its own global reductions (RED.ADD.F64 / RED.ADD.U64 at L2).  No loops, no match_any;
// exact for the integer counts, order-independent up to rounding for the fp64 sums.
__device__ __forceinline__ void image_add(const ImageDev& im, int bin, double w_flux, double w_real,
                                          double w_imag, unsigned count) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    // Every ray of the warp that is kept lands on ONE pixel (a focused system: the rays of a warp
    // are the pupil samples of one field point): plain butterfly sums, no bin compares or selects,
    // one lane issues the reductions.  REDUX gives the largest bin; dropped rays carry weight zero.
    const int top = __reduce_max_sync(full, bin);
    if (top < 0) return;
    if (__all_sync(full, bin == top || bin < 0)) {
        const bool keep = bin >= 0;
        w_flux = warp_sum(keep ? w_flux : 0.0);
        w_real = warp_sum(keep ? w_real : 0.0);
        if (im.moment_imag) w_imag = warp_sum(keep ? w_imag : 0.0);
        count = __reduce_add_sync(full, keep ? count : 0u);
        if (lane == 0) {
            if (im.flux) atomicAdd(im.flux + top, w_flux);
            if (im.moment_real) atomicAdd(im.moment_real + top, w_real);
            if (im.moment_imag) atomicAdd(im.moment_imag + top, w_imag);
            if (im.counts) atomicAdd(im.counts + top, (unsigned long long)count);
        }
        return;
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int pb = __shfl_xor_sync(full, bin, o);
        const double pf = __shfl_xor_sync(full, w_flux, o);
        const double pr = __shfl_xor_sync(full, w_real, o);
        const unsigned pc = __shfl_xor_sync(full, count, o);
        const bool same = (pb == bin) && (bin >= 0);
        const bool lower = (lane & o) == 0;
        if (same && lower) {
            w_flux += pf;
            w_real += pr;
            count += pc;
        }
        if (im.moment_imag) {
            const double pi = __shfl_xor_sync(full, w_imag, o);
            if (same && lower) w_imag += pi;
        }
        if (same && !lower) bin = -1;
    }
    if (bin >= 0) {
        if (im.flux) atomicAdd(im.flux + bin, w_flux);
        if (im.moment_real) atomicAdd(im.moment_real + bin, w_real);
        if (im.moment_imag) atomicAdd(im.moment_imag + bin, w_imag);
        if (im.counts) atomicAdd(im.counts + bin, (unsigned long long)count);
    }
}

// Bin index of one ray given in sensor-local coordinates, or -1 (dropped / vignetted).
// flux = intensity * where (optika/sensors/_sensors.py:139): vignetted rays carry weight
// zero and are not counted.
// (an image has fewer than 2^31 bins, checked in fill_image: a 32-bit index halves the shuffles)
__device__ __forceinline__ int image_bin_index(const ImageDev& im, const ImageGuess& g, bool valid,
                                                     double wavelength, double x, double y, bool unvignetted) {
    if (!valid || !unvignetted) return -1;
    const int iw = find_bin_search(im.e_w, im.n_w, wavelength, g.w0, g.w1);
    const int ix = find_bin_guess(im.e_x, im.n_x, x, g.x0, g.x1, g.inv_dx);
    const int iy = find_bin_guess(im.e_y, im.n_y, y, g.y0, g.y1, g.inv_dy);
    if (iw < 0 || ix < 0 || iy < 0) return -1;
    return (iw * im.n_x + ix) * im.n_y + iy;
}

// Bin one ray given in sensor-local coordinates.  All lanes of the warp call this.
__device__ __forceinline__ void image_bin_ray(const ImageDev& im, const ImageGuess& g, bool valid, double wavelength, double x,
                                              double y, double cos_real, double cos_imag, double intensity,
                                              bool unvignetted) {
    const int bin = image_bin_index(im, g, valid, wavelength, x, y, unvignetted);
    image_add(im, bin, intensity, intensity * cos_real, intensity * cos_imag, 1u);
}

}  // namespace optk
