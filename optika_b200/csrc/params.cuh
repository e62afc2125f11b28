// Kernel parameter blocks shared by the launchers (trace.cu, multilayer.cu) and
// the C ABI glue (api.cu).  Passed by value as __grid_constant__ kernel
// parameters: they live in the constant bank and are broadcast to the warp.
#pragma once
#include "common.cuh"
#include "bin.cuh"

namespace optk {

struct TraceParams {
    optk_rays_in_t in;
    optk_rays_out_t out;
    FastDiv div[OPTK_MAX_AXES];
    long long n_rays;
    long long index_offset;  // added to the thread index before decomposing it into grid indices
    long long accumulate_stride;
    long long prefetch_distance;  // rays; 0 = off.  L2 prefetch of the rays a later CTA will load
    // Broadcast (strided) inputs are addressed in two levels: the leading `n_axes - n_inner_axes`
    // axes are fixed per CTA (offsets computed once per CTA), the trailing ones per thread.
    long long inner_size;       // product of the trailing n_inner_axes dims
    long long tiles_per_outer;  // CTAs per index of the leading axes
    int32_t n_inner_axes;
    int32_t flat_index;  // host slab path: n_rays is a slab of the grid, decompose every axis per thread
    // broadcast views whose offsets fit 32 bits (the usual case: small separable arrays):
    // element strides of the fields (and the mask, index OPTK_NUM_FIELDS) as int32
    int32_t offsets32;
    int32_t cta_count;  // CTAs that hold rays (see cta_rows)
    int32_t stride32[OPTK_NUM_FIELDS + 1][OPTK_MAX_AXES];
    int32_t n_surf;
    int32_t accumulate;
    int32_t dense_in;  // every input is a dense array indexed by the thread index
    int32_t has_image;
    int32_t has_frame;
    // Fused image launches visit the ray grid in a strided order: CTA b works on tile
    // (b mod 512) * cta_rows + b / 512, so the ~450 CTAs resident at any time are spread over the
    // whole grid instead of sitting on ~20 neighbouring field points whose rays all add to the same
    // few pixels (same-address reductions serialise in L2).  0 = identity order.
    int32_t cta_rows;
    ImageDev image;
    optk_affine_t frame;
    optk_trace_stats_t* stats;
    // on-device ray generator (optk_trace_grid): `in` then only carries n_axes = 5 and
    // dims = grid.count (the sub-box), addressed in two levels like a broadcast view
    int32_t from_grid;
    int32_t has_out;  // any output pointer is set (fused image calls usually write no rays)
    int32_t relative_done;  // launch_trace has composed the frames of consecutive surfaces (once per parameter block)
    int32_t pad_relative;
    FastDiv div_tiles;  // divisor tiles_per_outer
    optk_grid_t grid;
    unsigned long long cell_stride[5];  // C-order strides of the whole grid n[] (Philox counter)
    optk_surface_t surf[OPTK_MAX_SURFACES];
};

#ifndef __CUDACC_RTC__
// run-time specialised kernels (jit.cu)
struct JitVariant {
    int dense, vec, image, grid, minb;
    int groups = 0;  // optk_image_t::group_size: accumulate per group of rays instead of per pixel
    // launch-wide facts compiled into the kernel (bit 8 = "defined"): OPTK_IMAGE_* bits of bin.cuh,
    // OPTK_GRID_* bits of trace_impl.cuh
    int image_flags = 0, grid_flags = 0;
};
void* jit_kernel(const TraceParams& P, const JitVariant& v);  // CUfunction or nullptr
int jit_launch(void* function, const TraceParams& P, unsigned grid, cudaStream_t stream);
void jit_set_mode(int mode);  // -1 automatic (long launches), 0 off, 1 always
long long jit_compiled_count();
int launch_trace(const TraceParams& P, cudaStream_t stream);
int launch_stop_newton(const TraceParams& P, const optk_stop_problem_t& problem, int sag_slot, long long n,
                       const double* wavelength, const double* const fixed[3], const double* const target[2], double* x,
                       double* y, double* z, unsigned int* n_unconverged, cudaStream_t stream);
int launch_trace_tma(const TraceParams& P, cudaStream_t stream);  // full tiles of tma_tile_rays() rays only
int tma_tile_rays();
int launch_bin(long long n_rays, const double* wavelength, const double* x, const double* y, const double* dz,
               const double* intensity, const uint8_t* unvignetted, const ImageDev& im, cudaStream_t stream);
#endif

struct LayerDev {
    const double* n_re;
    const double* n_im;
    const double* thickness;
    const double* width;
    long long n_stride[OPTK_ML_MAX_AXES];
    long long t_stride[OPTK_ML_MAX_AXES];
    long long w_stride[OPTK_ML_MAX_AXES];
    int32_t profile_kind;
    // fast addressing: the single grid axis each array varies along (-1: none, -2: several, use strides)
    int8_t n_axis, t_axis, w_axis, pad;
};

struct MultilayerParams {
    optk_ml_input_t in;
    FastDiv div[OPTK_ML_MAX_AXES];
    long long n_eval;
    const LayerDev* layers;  // device, n_layers entries; the last one is the substrate
    int32_t n_layers;
    int32_t n_segments;
    optk_ml_segment_t segments[32];
    double* r_s;
    double* r_p;
    double* t_s;
    double* t_p;
    // An axis along which ONLY layer thicknesses vary (-1: none): the Fresnel terms are shared by
    // all its indices, so one thread evaluates several of them (cross-configuration reuse).
    int32_t reuse_axis;
    int32_t pad;
};

#ifndef __CUDACC_RTC__
int launch_multilayer(const MultilayerParams& P, cudaStream_t stream);
int launch_interp(long long n, const double* x, int m, const double* xp, const double* fp_re, const double* fp_im,
                  double* out_re, double* out_im, cudaStream_t stream);
int launch_apply_efficiency(long long n, double* intensity, const double* e_s, const double* e_p, cudaStream_t stream);
int launch_debug_math(int op, long long n, const double* a, const double* b, double* out, cudaStream_t stream);
int launch_reduce_groups(long long n_groups, long long n_inner, const double* x, const double* y, const double* intensity,
                         const uint8_t* unvignetted, double* sum_intensity, double* sum_x, double* sum_y,
                         unsigned long long* count, double* sum_x_all, double* sum_y_all, cudaStream_t stream);
int launch_electrons(int n_plane, int n_x, int n_y, const optk_ccd_plane_t* planes_device, const long long* photons,
                     unsigned long long* electrons, int wrap, unsigned long long seed, cudaStream_t stream);
int measure_fp64_peak(double* flops, cudaStream_t stream);
int measure_soa_copy(long long n_rays, double* gbytes_per_second, cudaStream_t stream);

#endif

}  // namespace optk
