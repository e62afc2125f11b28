// Kernel 1: the fused sequential raytrace (sm_100a, fp64): generic-path instantiations
// (stage masks, caller normals, sag transformations: one ray per thread), the host
// launcher, and the standalone binning kernel.  The device code is in trace_impl.cuh.
#include "trace_impl.cuh"

namespace optk {

trace_kernel_t select_generic_kernel(bool dense, bool acc, bool image) {
#define OPTK_PICK(D, A, I) \
    if (dense == D && acc == A && image == I) return (trace_kernel_t)trace_kernel<4, 1, false, D, false, A, I>;
    OPTK_PICK(true, false, false)
    OPTK_PICK(true, true, false)
    OPTK_PICK(true, false, true)
    OPTK_PICK(true, true, true)
    OPTK_PICK(false, false, false)
    OPTK_PICK(false, true, false)
    OPTK_PICK(false, false, true)
    OPTK_PICK(false, true, true)
#undef OPTK_PICK
    return nullptr;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------------------
// relative frames: one affine per transition of the walk instead of "local -> global" after a surface and
// "global -> local" before the next (AbstractSurface.propagate_rays, optika/surfaces.py:141-142, 195-196,
// composed on the host in double precision; results differ from the two-step form by rounding, 1e-16)
// ---------------------------------------------------------------------------
static void relative_frames(TraceParams& Q) {
    auto frame_of = [](const optk_surface_t& S, double (&r)[9], double (&t)[3]) {
        static const double eye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        const bool has = (S.flags & OPTK_F_TRANSFORM) != 0;
        for (int i = 0; i < 9; ++i) r[i] = has ? S.transform.r[i] : eye[i];
        for (int i = 0; i < 3; ++i) t[i] = has ? S.transform.t[i] : 0.0;
    };
    for (int k = 1; k < Q.n_surf; ++k) {
        optk_surface_t& A = Q.surf[k - 1];
        optk_surface_t& B = Q.surf[k];
        if (A.flags & OPTK_F_LOCAL_OUT) continue;  // the caller wants A's rays local: leave the pair alone
        if (!((A.flags | B.flags) & OPTK_F_TRANSFORM)) continue;  // both in the global frame already
        double ra[9], ta[3], rb[9], tb[3];
        frame_of(A, ra, ta);
        frame_of(B, rb, tb);
        // x_B = R_B^T (R_A x_A + t_A - t_B)
        bool same = true;
        for (int i = 0; i < 9; ++i) same = same && ra[i] == rb[i];
        bool same_t = true;
        for (int i = 0; i < 3; ++i) same_t = same_t && ta[i] == tb[i];
        optk_affine_t rel;
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) {
                double v = 0.0;
                for (int m = 0; m < 3; ++m) v += rb[3 * m + i] * ra[3 * m + j];
                rel.r[3 * i + j] = same ? (i == j ? 1.0 : 0.0) : v;
            }
            double v = 0.0;
            for (int m = 0; m < 3; ++m) v += rb[3 * m + i] * (ta[m] - tb[m]);
            rel.t[i] = v;
        }
        A.flags |= OPTK_F_LOCAL_OUT;
        B.flags |= OPTK_F_RELATIVE_IN;
        if (same && same_t) B.flags |= OPTK_F_RELATIVE_IDENTITY;
        else if (same) {
            bool identity = true;  // a shared rotation other than the identity still rotates the offset: general case
            for (int i = 0; i < 9; ++i) identity = identity && rb[i] == (i % 4 == 0 ? 1.0 : 0.0);
            if (identity) B.flags |= OPTK_F_RELATIVE_TRANSLATION;
        }
        B.sag_transform = rel;
    }
}

// ---------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------
int launch_trace(const TraceParams& P, cudaStream_t stream) {
    if (P.n_rays <= 0) return OPTK_OK;
    const bool from_grid = P.from_grid != 0;
    bool full = P.in.normal[0] == nullptr && P.out.cos_incidence == nullptr;
    for (int s = 0; s < P.n_surf; ++s)
        full = full && (P.surf[s].stages == OPTK_STAGE_ALL) && !(P.surf[s].flags & OPTK_F_SAG_TRANSFORM) &&
               P.surf[s].material_kind <= OPTK_MAT_GLASS;
    bool efficiency = false;  // measured mirrors / rulings, groove profiles: the EFF instantiations
    for (int s = 0; s < P.n_surf; ++s)
        efficiency = efficiency || P.surf[s].material_efficiency != OPTK_EFF_UNIT ||
                     P.surf[s].ruling_profile != OPTK_PROFILE_IDEAL;
    // curvilinear grids with non-unit efficiencies are rare enough to share the generic kernels
    const bool curvilinear = from_grid && (P.grid.field_2d || P.grid.pupil_2d || P.grid.chromatic);
    if (curvilinear && efficiency) full = false;
    const bool dense = P.dense_in != 0 && !from_grid, acc = P.accumulate != 0, image = P.has_image != 0;
    // 128-bit path: dense inputs, every array 16-byte aligned, even accumulate stride
    bool vec = full && dense && !from_grid && !efficiency && (P.accumulate_stride % 2 == 0);
    for (int f = 0; f < OPTK_NUM_FIELDS && vec; ++f)
        vec = aligned16(P.in.field[f]) && aligned16(P.out.field[f]);
    if (vec && P.in.unvignetted) vec = (reinterpret_cast<uintptr_t>(P.in.unvignetted) & 1u) == 0;
    if (vec && P.out.unvignetted) vec = (reinterpret_cast<uintptr_t>(P.out.unvignetted) & 1u) == 0;
    const int rays_per_thread = full ? 2 : 1;
    const int block = 256;
    TraceParams& Q = const_cast<TraceParams&>(P);
    static const int relative_mode = [] {
        const char* e = getenv("OPTK_TRACE_RELATIVE");
        return e ? atoi(e) : 1;
    }();
    // (idempotent: a surface that already arrives relative is skipped, e.g. the tail launch of the pipeline split)
    if (full && !acc && relative_mode && P.n_surf > 1 && !(P.surf[1].flags & OPTK_F_RELATIVE_IN) && !P.relative_done) {
        relative_frames(Q);
        Q.relative_done = 1;
    }
    // HBM-streaming case (dense in, dense out, nothing else): the persistent bulk-copy pipeline
    // takes the full 512-ray tiles, the ordinary kernel the remainder
    // The pipeline runs 16 warps per SM (114 registers, no spills), the direct-load kernel 24:
    // the pipeline wins while the walk is HBM-bound (cfg 2: 2.8 vs 3.2 ms per 1e8 rays) and
    // loses once it is FP64/issue-bound (cfg 1, six surfaces: 4.5 vs 4.2 ms; cfg 3, toroid:
    // 9.7 vs 8.2 ms).  Rough per-ray cost in units of a flat surface decides; OPTK_TRACE_TMA
    // = 0 / 1 forces the choice.
    static const int tma_mode = [] {
        const char* e = getenv("OPTK_TRACE_TMA");
        return e ? (atoi(e) != 0 ? 1 : 0) : -1;
    }();
    // rough per-ray cost of the surface list in units of a flat surface
    double cost = 0.0;
    bool out_of_line = false;  // a sag or ruling kind that is a call in the streamlined kernels
    for (int s = 0; s < P.n_surf; ++s) {
        const optk_surface_t& S = P.surf[s];
        out_of_line = out_of_line || S.ruling_kind > OPTK_RULING_CONSTANT ||
                      !(S.sag_kind == OPTK_SAG_FLAT || S.sag_kind == OPTK_SAG_SPHERICAL || S.sag_kind == OPTK_SAG_PARABOLIC);
        cost += S.sag_kind == OPTK_SAG_FLAT ? 1.0
                : (S.sag_kind == OPTK_SAG_SPHERICAL || S.sag_kind == OPTK_SAG_PARABOLIC) ? 1.3
                : S.sag_kind == OPTK_SAG_TOROIDAL ? 4.0 : 2.0;
        if (S.ruling_kind > OPTK_RULING_CONSTANT) cost += 1.0;
        if (S.material_kind == OPTK_MAT_GLASS) cost += 0.5;
    }
    const bool heavy = cost > 4.5;
    static const int jit_minb = [] {
        const char* e = getenv("OPTK_JIT_MINB");  // experiments: resident CTAs per SM the compiled kernel is built for
        return e ? atoi(e) : 0;
    }();
    // A kernel compiled for this surface list (jit.cu) streams faster than the pipeline: its walk is
    // short enough for 24 resident warps to cover HBM latency (cfg 2: 2.41 ms per 1e8 rays against
    // 2.74 for the pipeline and 3.18 for the table-driven direct-load kernel).
    bool specialised = false;
    if (tma_mode < 0 && !heavy && full && vec && !acc && !image && !from_grid) {
        const JitVariant variant = {1, 1, 0, 0, jit_minb > 0 ? jit_minb : 3};
        specialised = jit_kernel(P, variant) != nullptr;
    }
    const bool use_tma = tma_mode == 1 || (tma_mode < 0 && !heavy && !specialised);
    if (use_tma && vec && !acc && !image && !from_grid && P.n_rays >= 64LL * tma_tile_rays()) {
        // bulk copies need 16-byte aligned sources (the fields are, `vec`; the mask may not be)
        bool all_out = P.out.unvignetted != nullptr && aligned16(P.in.unvignetted);
        for (int f = 0; f < OPTK_NUM_FIELDS; ++f) all_out = all_out && P.out.field[f] != nullptr;
        if (all_out) {
            const long long tile = tma_tile_rays();
            const long long n_main = P.n_rays / tile * tile, n_tail = P.n_rays - n_main;
            const long long n_all = P.n_rays;
            Q.n_rays = n_main;
            int rc = launch_trace_tma(P, stream);
            Q.n_rays = n_all;
            if (rc || n_tail == 0) return rc;
            static thread_local TraceParams T;
            T = P;
            T.n_rays = n_tail;
            for (int f = 0; f < OPTK_NUM_FIELDS; ++f) {
                T.in.field[f] += n_main;
                T.out.field[f] += n_main;
            }
            if (T.in.unvignetted) T.in.unvignetted += n_main;
            T.out.unvignetted += n_main;
            T.in.dims[0] = n_tail;  // dense: one flat axis is all the kernel looks at
            return launch_trace(T, stream);
        }
    }
    long long grid;
    if (dense || from_grid) {
        const long long threads = (P.n_rays + rays_per_thread - 1) / rays_per_thread;
        grid = (threads + block - 1) / block;
        Q.n_inner_axes = 0;
        Q.inner_size = P.n_rays;
        Q.tiles_per_outer = grid;
    } else {
        // trailing axes until a CTA-sized tile wastes little; a slab offset (host path) or a
        // normal array keeps the flat scheme (every axis per thread)
        int n_inner = 0;
        long long inner = 1;
        while (n_inner < P.in.n_axes && (inner < 16LL * block * rays_per_thread || P.flat_index)) {
            inner *= P.in.dims[P.in.n_axes - 1 - n_inner];
            ++n_inner;
        }
        if (P.flat_index) inner = P.n_rays;  // slab of a larger grid: one "outer" index
        const long long per_tile = (long long)block * rays_per_thread;
        Q.n_inner_axes = n_inner;
        Q.inner_size = inner;
        Q.tiles_per_outer = (inner + per_tile - 1) / per_tile;
        grid = (inner > 0 ? P.n_rays / inner : 0) * Q.tiles_per_outer;
    }
    Q.div_tiles = make_fastdiv((uint32_t)(Q.tiles_per_outer > 0 ? Q.tiles_per_outer : 1));
    Q.has_out = (P.out.unvignetted != nullptr || P.out.cos_incidence != nullptr) ? 1 : 0;
    for (int f = 0; f < OPTK_NUM_FIELDS; ++f)
        if (P.out.field[f]) Q.has_out = 1;
    if (grid > 0x7fffffffLL) {
        set_error("optk_trace: too many rays for one launch (%lld)", P.n_rays);
        return OPTK_ERR_INVALID;
    }
    trace_kernel_t kernel;
    static const int heavy_mode = [] {
        const char* e = getenv("OPTK_TRACE_HEAVY");
        return e ? (atoi(e) != 0 ? 1 : 0) : -1;
    }();
    // (with calls in the walk the extra registers do not pay: cfg 3 fused 9.8 -> 11.1 ms)
    // Round 2: after the diet of the binning and generator code the 24-warp build wins for long
    // surface lists too (cfg 5, 1.07e10 rays: 365 ms against 379 with 16 warps and 110 registers), so the 16-warp
    // build is only taken on request (OPTK_TRACE_HEAVY=1).
    (void)out_of_line;
    const bool use_heavy = full && from_grid && image && heavy_mode == 1;
    if (curvilinear)
        kernel = select_grid_kernel(full, acc, image, true);
    else if (full && efficiency)
        kernel = select_efficiency_kernel(from_grid, dense, acc, image);
    else if (use_heavy)
        kernel = select_heavy_kernel(from_grid, acc, image);
    else if (from_grid)
        kernel = select_grid_kernel(full, acc, image, false);
    else if (full)
        kernel = select_full_kernel(dense, vec, acc, image);
    else
        kernel = select_generic_kernel(dense, acc, image);
    static const int prefetch_waves = [] {
        const char* e = getenv("OPTK_TRACE_PREFETCH");
        return e ? atoi(e) : 0;  // measured: no gain once 24 warps per SM are resident
    }();
    Q.prefetch_distance = 0;
    Q.offsets32 = 0;
    if (!dense) {
        bool fits = true;
        for (int f = 0; f <= OPTK_NUM_FIELDS && fits; ++f) {
            const int64_t* st = f < OPTK_NUM_FIELDS ? P.in.stride[f] : P.in.mask_stride;
            long long extent = 0;
            for (int a = 0; a < P.in.n_axes; ++a) {
                if (st[a] < 0) fits = false;
                extent += (P.in.dims[a] - 1) * st[a];
                Q.stride32[f][a] = (int32_t)st[a];
            }
            if (extent >= 0x7fffffffLL) fits = false;
        }
        Q.offsets32 = fits ? 1 : 0;
    }
    if (dense && prefetch_waves > 0) {
        int device = 0, sms = 0, ctas = 0;
        OPTK_CUDA(cudaGetDevice(&device));
        OPTK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        OPTK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, (const void*)kernel, block, 0));
        Q.prefetch_distance = (long long)prefetch_waves * sms * ctas * block * rays_per_thread;
    }
    // fused image launches without a dense input stream: strided visiting order (params.cuh)
    Q.cta_rows = 0;
    Q.cta_count = (int32_t)grid;
    static const int spread_mode = [] {
        const char* e = getenv("OPTK_TRACE_SPREAD");
        return e ? atoi(e) : 1;
    }();
    if (image && !dense && spread_mode && grid > 512) {
        const long long rows = (grid + 511) / 512;
        if (rows * 512 <= 0x7fffffffLL) {
            Q.cta_rows = (int32_t)rows;
            grid = rows * 512;
        }
    }
    // a kernel compiled for exactly this surface list, when the launch is long enough to pay for it
    if (full && !acc) {
        JitVariant variant = {dense ? 1 : 0, vec ? 1 : 0, image ? 1 : 0, from_grid ? (curvilinear ? 2 : 1) : 0,
                              jit_minb > 0 ? jit_minb : (use_heavy ? 2 : 3)};
        variant.groups = (image && P.image.group) ? 1 : 0;
        if (image && !P.image.group)
            variant.image_flags = 0x100 | (P.image.has_range ? 1 << OPTK_IMAGE_HAS_RANGE : 0) |
                                  ((P.image.uniform & 1) ? 1 << OPTK_IMAGE_UNIFORM_X : 0) |
                                  ((P.image.uniform & 2) ? 1 << OPTK_IMAGE_UNIFORM_Y : 0) |
                                  (P.image.n_w == 1 ? 1 << OPTK_IMAGE_ONE_WAVELENGTH : 0) |
                                  (P.image.counts ? 1 << OPTK_IMAGE_COUNTS : 0) |
                                  (P.image.moment_real ? 1 << OPTK_IMAGE_MOMENT_REAL : 0) |
                                  (P.image.flux ? 1 << OPTK_IMAGE_FLUX : 0) |
                                  (P.image.moment_imag ? 1 << OPTK_IMAGE_MOMENT_IMAG : 0);
        if (from_grid)
            variant.grid_flags = 0x100 | (P.grid.at_infinity ? 1 << OPTK_GRID_AT_INFINITY : 0) |
                                 ((P.grid.angular_cells[0] && !curvilinear) ? 1 << OPTK_GRID_PACKED : 0) |
                                 (P.grid.jitter ? 1 << OPTK_GRID_JITTER : 0) |
                                 (P.grid.has_frame ? 1 << OPTK_GRID_FRAME : 0) |
                                 (P.grid.weight_scene ? 1 << OPTK_GRID_WEIGHT_SCENE : 0) |
                                 (P.grid.weight_pupil ? 1 << OPTK_GRID_WEIGHT_PUPIL : 0);
        if (void* function = jit_kernel(P, variant)) return jit_launch(function, P, (unsigned)grid, stream);
    }
    if (image && P.image.group) {
        set_error("group accumulators (optk_image_t.group_size) are served by the run-time compiled kernels only "
                  "(full operator, no accumulate, NVRTC available, OPTK_JIT != 0); use optk_reduce_groups on the traced rays");
        return OPTK_ERR_UNSUPPORTED;
    }
    void* args[] = {(void*)&P};
    OPTK_CUDA(cudaLaunchKernel((const void*)kernel, dim3((unsigned)grid), dim3(block), args, 0, stream));
    return OPTK_OK;
}

// ---------------------------------------------------------------------------
// standalone binning kernel (kernel 2) for rays already in sensor coordinates
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bin_kernel(long long n_rays, const double* __restrict__ wavelength, const double* __restrict__ x,
           const double* __restrict__ y, const double* __restrict__ dz, const double* __restrict__ intensity,
           const uint8_t* __restrict__ unvignetted, const __grid_constant__ ImageDev im, int cta_rows, int cta_count) {
    // strided visiting order (TraceParams::cta_rows): rays arrive field by field, and the CTAs resident
    // at one time would otherwise all add to the same few pixels
    unsigned block = blockIdx.x;
    if (cta_rows) {
        block = (block & 511u) * (unsigned)cta_rows + (block >> 9);
        if (block >= (unsigned)cta_count) return;
    }
    __shared__ ImageGuess guess;
    image_guess_init(im, &guess);
    const long long i = (long long)block * blockDim.x + threadIdx.x;
    const bool valid = i < n_rays;
    double w = 0, px = 0, py = 0, c = 1.0, in = 0;
    bool unv = false;
    if (valid) {
        w = __ldg(wavelength + i);
        px = __ldg(x + i);
        py = __ldg(y + i);
        c = dz ? __ldg(dz + i) : 1.0;
        in = intensity ? __ldg(intensity + i) : 1.0;
        unv = unvignetted ? (__ldg(unvignetted + i) != 0) : true;
    }
    image_bin_ray(im, guess, valid, w, px, py, c, 0.0, in, unv);
}

int launch_bin(long long n_rays, const double* wavelength, const double* x, const double* y, const double* dz,
               const double* intensity, const uint8_t* unvignetted, const ImageDev& im, cudaStream_t stream) {
    if (n_rays <= 0) return OPTK_OK;
    const int block = 256;
    long long grid = (n_rays + block - 1) / block;
    const long long rows = (grid + 511) / 512;
    if (rows * 512 > 0x7fffffffLL) {
        set_error("optk_bin: too many rays for one launch (%lld)", n_rays);
        return OPTK_ERR_INVALID;
    }
    int cta_rows = 0;
    const int cta_count = (int)grid;
    if (grid > 512) {
        cta_rows = (int)rows;
        grid = rows * 512;
    }
    bin_kernel<<<(unsigned)grid, block, 0, stream>>>(n_rays, wavelength, x, y, dz, intensity, unvignetted, im, cta_rows,
                                                      cta_count);
    OPTK_CUDA(cudaGetLastError());
    return OPTK_OK;
}

}  // namespace optk
