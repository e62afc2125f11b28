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

// What a launch knows about its image before the first ray (which planes exist, uniform edges, one
// spectral bin) is a run-time test of a kernel parameter in the table-driven kernels and a compile-time
// constant in the run-time specialised ones (jit.cu defines OPTK_JIT_IMAGE_FLAGS for the variant).
#define OPTK_IMAGE_HAS_RANGE 0
#define OPTK_IMAGE_UNIFORM_X 1
#define OPTK_IMAGE_UNIFORM_Y 2
#define OPTK_IMAGE_ONE_WAVELENGTH 3
#define OPTK_IMAGE_COUNTS 4
#define OPTK_IMAGE_MOMENT_REAL 5
#define OPTK_IMAGE_FLUX 6
#define OPTK_IMAGE_MOMENT_IMAG 7
#ifdef OPTK_JIT_IMAGE_FLAGS
#define OPTK_IMAGE_FLAG(bit, runtime) ((((OPTK_JIT_IMAGE_FLAGS) >> (bit)) & 1) != 0)
#else
#define OPTK_IMAGE_FLAG(bit, runtime) (runtime)
#endif

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
    int32_t uniform;    // optk_image_t::uniform_edges: bit 0 the x edges, bit 1 the y edges are a linspace
    double range[6];    // first / last edge of wavelength, x, y
    double inv_dx, inv_dy;  // n / (last - first) of the pixel axes, computed on the host when has_range
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

// The same from the constant bank alone (has_range: the host computed the reciprocal widths too).
__device__ __forceinline__ ImageGuess image_guess_const(const ImageDev& im) {
    ImageGuess g;
    g.w0 = im.range[0]; g.w1 = im.range[1];
    g.x0 = im.range[2]; g.x1 = im.range[3];
    g.y0 = im.range[4]; g.y1 = im.range[5];
    g.inv_dx = im.inv_dx;
    g.inv_dy = im.inv_dy;
    return g;
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

// The same, out of line: the rarely taken exact path behind find_bin_uniform.
static __device__ __noinline__ int find_bin_exact(const double* __restrict__ e, int n, double v, double e0, double e1,
                                                  double inv_d) {
    return find_bin_guess(e, n, v, e0, e1, inv_d);
}

// Edges the caller announces as uniform (optk_image_t::uniform_edges: every edge within 1e-9 of a bin width
// of first + i (last - first) / n, which is what linspace produces): u = (v - e0) n / (e1 - e0) is then the
// bin coordinate to ~1e-12, so floor(u) IS the bin unless u lies within 1e-6 of an integer -- only there
// (two rays in a million, and everything outside the range or NaN) are the exact edge VALUES consulted.
// No loads, no loops on the common path; same result as find_bin_guess for every sample.
__device__ __forceinline__ int find_bin_uniform(const double* __restrict__ e, int n, double v, double e0, double e1,
                                                double inv_d) {
    const double u = (v - e0) * inv_d;
    const int i = __double2int_rd(u);
    const double f = u - (double)i;
    if (fabs(f - 0.5) < 0.499999 && (unsigned)i < (unsigned)n) return i;
    return find_bin_exact(e, n, v, e0, e1, inv_d);
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
    const bool has_flux = OPTK_IMAGE_FLAG(OPTK_IMAGE_FLUX, im.flux != nullptr);
    const bool has_real = OPTK_IMAGE_FLAG(OPTK_IMAGE_MOMENT_REAL, im.moment_real != nullptr);
    const bool has_imag = OPTK_IMAGE_FLAG(OPTK_IMAGE_MOMENT_IMAG, im.moment_imag != nullptr);
    const bool has_counts = OPTK_IMAGE_FLAG(OPTK_IMAGE_COUNTS, im.counts != nullptr);
    // Every ray of the warp that is kept lands on ONE pixel (a focused system: the rays of a warp
    // are the pupil samples of one field point): plain butterfly sums, no bin compares or selects,
    // one lane issues the reductions.  REDUX gives the largest bin; dropped rays carry weight zero.
    const int top = __reduce_max_sync(full, bin);
    if (top < 0) return;
    const bool keep_lane = bin >= 0;
    if (__all_sync(full, bin == top || !keep_lane)) {
        const bool keep = bin >= 0;
        if (has_flux) w_flux = warp_sum(keep ? w_flux : 0.0);
        if (has_real) w_real = warp_sum(keep ? w_real : 0.0);
        if (has_imag) w_imag = warp_sum(keep ? w_imag : 0.0);
        if (has_counts) count = __reduce_add_sync(full, keep ? count : 0u);
        if (lane == 0) {
            if (has_flux) atomicAdd(im.flux + top, w_flux);
            if (has_real) atomicAdd(im.moment_real + top, w_real);
            if (has_imag) atomicAdd(im.moment_imag + top, w_imag);
            if (has_counts) atomicAdd(im.counts + top, (unsigned long long)count);
        }
        return;
    }
    // Would merging pay?  Lanes that are neighbours here are neighbours in the pupil, so the first
    // step of the network is a fair sample: when fewer than half of the kept rays share their pixel
    // with the neighbouring lane (stratified random samples of a cell that spans several pixels),
    // the five compare-and-merge rounds (~100 instructions) would retire almost nobody -- every kept
    // lane issues its own reductions instead.  The sums are the same either way.
    // (OPTK_BIN_DIRECT=0 in the environment of a process compiles this test out of its run-time
    // specialised kernels: A/B measurements)
#if !defined(OPTK_JIT_BIN_DIRECT) || OPTK_JIT_BIN_DIRECT
    {
        const int nb = __shfl_xor_sync(full, bin, 1);
        const unsigned paired = __ballot_sync(full, keep_lane && nb == bin);
        const unsigned kept = __ballot_sync(full, keep_lane);
        if (2 * __popc(paired) < __popc(kept)) {
            if (keep_lane) {
                if (has_flux) atomicAdd(im.flux + bin, w_flux);
                if (has_real) atomicAdd(im.moment_real + bin, w_real);
                if (has_imag) atomicAdd(im.moment_imag + bin, w_imag);
                if (has_counts) atomicAdd(im.counts + bin, (unsigned long long)count);
            }
            return;
        }
    }
#endif
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int pb = __shfl_xor_sync(full, bin, o);
        const bool same = (pb == bin) && (bin >= 0);
        const bool lower = (lane & o) == 0;
        if (has_flux) {
            const double pf = __shfl_xor_sync(full, w_flux, o);
            if (same && lower) w_flux += pf;
        }
        if (has_real) {
            const double pr = __shfl_xor_sync(full, w_real, o);
            if (same && lower) w_real += pr;
        }
        if (has_counts) {
            const unsigned pc = __shfl_xor_sync(full, count, o);
            if (same && lower) count += pc;
        }
        if (has_imag) {
            const double pi = __shfl_xor_sync(full, w_imag, o);
            if (same && lower) w_imag += pi;
        }
        if (same && !lower) bin = -1;
    }
    if (bin >= 0) {
        if (has_flux) atomicAdd(im.flux + bin, w_flux);
        if (has_real) atomicAdd(im.moment_real + bin, w_real);
        if (has_imag) atomicAdd(im.moment_imag + bin, w_imag);
        if (has_counts) atomicAdd(im.counts + bin, (unsigned long long)count);
    }
}

// Bin index of one ray given in sensor-local coordinates, or -1 (dropped / vignetted).
// flux = intensity * where (optika/sensors/_sensors.py:139): vignetted rays carry weight
// zero and are not counted.
// (an image has fewer than 2^31 bins, checked in fill_image: a 32-bit index halves the shuffles)
__device__ __forceinline__ int image_bin_index(const ImageDev& im, const ImageGuess& g, bool valid,
                                                     double wavelength, double x, double y, bool unvignetted) {
    if (!valid || !unvignetted) return -1;
    // one spectral bin (SequentialSystem.image with integrate=True): a range test, no search
    const int iw = OPTK_IMAGE_FLAG(OPTK_IMAGE_ONE_WAVELENGTH, im.n_w == 1)
                       ? ((wavelength >= g.w0 && wavelength <= g.w1) ? 0 : -1)
                       : find_bin_search(im.e_w, im.n_w, wavelength, g.w0, g.w1);
    // (edges not announced as uniform take the exact search out of line: sensors have linspace edges,
    // and the kernel body is instruction-cache bound)
    const int ix = OPTK_IMAGE_FLAG(OPTK_IMAGE_UNIFORM_X, im.uniform & 1) ? find_bin_uniform(im.e_x, im.n_x, x, g.x0, g.x1, g.inv_dx)
                                    : find_bin_exact(im.e_x, im.n_x, x, g.x0, g.x1, g.inv_dx);
    const int iy = OPTK_IMAGE_FLAG(OPTK_IMAGE_UNIFORM_Y, im.uniform & 2) ? find_bin_uniform(im.e_y, im.n_y, y, g.y0, g.y1, g.inv_dy)
                                    : find_bin_exact(im.e_y, im.n_y, y, g.y0, g.y1, g.inv_dy);
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
