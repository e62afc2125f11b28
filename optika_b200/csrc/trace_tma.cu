// Kernel 1, HBM-streaming variant: dense SoA rays in, dense SoA rays out, full operator.
// A persistent CTA per (SM, slot) pulls 512-ray tiles of the ten input arrays (+ mask) into
// shared memory with 1-D bulk copies (cp.async.bulk, the TMA engine) two tiles ahead of the
// arithmetic, signalled by mbarriers; the surface walk is the same two-rays-per-thread code
// as the other kernels (surface_full) and the results go straight from registers to HBM
// with 128-bit stores.  Loads therefore never stall a warp and never occupy its registers:
// 128 registers per thread (no spills), two CTAs per SM, 83 KB of input in flight per SM.
#include "trace_impl.cuh"

namespace optk {

namespace {

constexpr int STAGES = 2;

template <int TILE>
struct alignas(128) TmaStage {
    double field[OPTK_NUM_FIELDS][TILE];
    uint8_t mask[TILE];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion is
// counted on the mbarrier in bytes
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int TILE>
__device__ __forceinline__ void issue_tile(const TraceParams& P, TmaStage<TILE>* stage, uint64_t* bar, long long tile,
                                           bool has_mask) {
    const unsigned bytes = OPTK_NUM_FIELDS * TILE * 8 + (has_mask ? TILE : 0);
    // the stage was last read through the generic proxy; order those reads before the
    // asynchronous-proxy writes of the copy engine
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(bar, bytes);
    const long long first = tile * TILE;
#pragma unroll
    for (int f = 0; f < OPTK_NUM_FIELDS; ++f) bulk_load(stage->field[f], P.in.field[f] + first, TILE * 8, bar);
    if (has_mask) bulk_load(stage->mask, P.in.unvignetted + first, TILE, bar);
}

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) trace_kernel_tma(const __grid_constant__ TraceParams P) {
    constexpr int TILE = 2 * THREADS;  // two rays per thread
    typedef TmaStage<TILE> Stage;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Stage* stages = reinterpret_cast<Stage*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + STAGES * sizeof(Stage));
    const long long n_tiles = P.n_rays / TILE;
    const bool has_mask = P.in.unvignetted != nullptr;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    long long tile = blockIdx.x;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            const long long t = tile + (long long)s * gridDim.x;
            if (t < n_tiles) issue_tile(P, &stages[s], &full[s], t, has_mask);
        }
    }

    unsigned long long n_unvignetted = 0;
    unsigned newton_iterations = 0;
    for (int it = 0; tile < n_tiles; tile += gridDim.x, ++it) {
        const int s = it % STAGES;
        const unsigned parity = (it / STAGES) & 1;
        mbar_wait(&full[s], parity);

        // the two rays of this thread: one 128-bit shared load per field
        Ray r[2];
        const Stage& st = stages[s];
        const int j = 2 * threadIdx.x;
        {
            double2 v;
            v = *reinterpret_cast<const double2*>(&st.field[OPTK_WAVELENGTH][j]); r[0].w = v.x; r[1].w = v.y;
            v = *reinterpret_cast<const double2*>(&st.field[OPTK_PX][j]); r[0].px = v.x; r[1].px = v.y;
            v = *reinterpret_cast<const double2*>(&st.field[OPTK_PY][j]); r[0].py = v.x; r[1].py = v.y;
            v = *reinterpret_cast<const double2*>(&st.field[OPTK_PZ][j]); r[0].pz = v.x; r[1].pz = v.y;
            v = *reinterpret_cast<const double2*>(&st.field[OPTK_DX][j]); r[0].dx = v.x; r[1].dx = v.y;
            v = *reinterpret_cast<const double2*>(&st.field[OPTK_DY][j]); r[0].dy = v.x; r[1].dy = v.y;
            v = *reinterpret_cast<const double2*>(&st.field[OPTK_DZ][j]); r[0].dz = v.x; r[1].dz = v.y;
            v = *reinterpret_cast<const double2*>(&st.field[OPTK_INTENSITY][j]); r[0].intensity = v.x; r[1].intensity = v.y;
            v = *reinterpret_cast<const double2*>(&st.field[OPTK_ATTENUATION][j]); r[0].att = v.x; r[1].att = v.y;
            v = *reinterpret_cast<const double2*>(&st.field[OPTK_INDEX_REFRACTION][j]); r[0].n = v.x; r[1].n = v.y;
            if (has_mask) {
                const uchar2 m = *reinterpret_cast<const uchar2*>(&st.mask[j]);
                r[0].unv = m.x != 0;
                r[1].unv = m.y != 0;
            } else {
                r[0].unv = r[1].unv = true;
            }
        }
        __syncthreads();  // every thread has its rays in registers: the stage is free
        if (threadIdx.x == 0) {
            const long long next = tile + (long long)STAGES * gridDim.x;
            if (next < n_tiles) issue_tile(P, &stages[s], &full[s], next, has_mask);
        }

        WalkState state = {(r[0].att != 0.0) || (r[1].att != 0.0), false, false};
        for (int k = 0; k < P.n_surf; ++k) surface_full<2>(P.surf[k], r, newton_iterations, state);
        store_rays_vec(P.out, tile * TILE + j, r);
        if (P.stats) n_unvignetted += (r[0].unv ? 1 : 0) + (r[1].unv ? 1 : 0);
    }

    if (P.stats) {
        const unsigned full_mask = 0xffffffffu;
        const unsigned n_unv = __reduce_add_sync(full_mask, (unsigned)n_unvignetted);
        const unsigned n_it = __reduce_add_sync(full_mask, newton_iterations);
        if ((threadIdx.x & 31) == 0) {
            if (n_unv) atomicAdd(&P.stats->n_unvignetted, (unsigned long long)n_unv);
            if (n_it) atomicAdd(&P.stats->n_newton_iterations, (unsigned long long)n_it);
        }
        if (blockIdx.x == 0 && threadIdx.x == 0)
            atomicAdd(&P.stats->n_rays, (unsigned long long)(n_tiles * TILE));
    }
}

template <int THREADS, int MINB>
int launch_shape(const TraceParams& P, cudaStream_t stream) {
    constexpr int TILE = 2 * THREADS;
    const void* kernel = (const void*)trace_kernel_tma<THREADS, MINB>;
    static int ctas_per_device = 0;
    const size_t smem = STAGES * sizeof(TmaStage<TILE>) + STAGES * sizeof(uint64_t);
    if (!ctas_per_device) {
        int device = 0, sms = 0;
        OPTK_CUDA(cudaGetDevice(&device));
        OPTK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        OPTK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        OPTK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, THREADS, smem));
        if (per_sm < 1) per_sm = 1;
        ctas_per_device = sms * per_sm;
    }
    const long long n_tiles = P.n_rays / TILE;
    if (n_tiles == 0) return OPTK_OK;
    const unsigned grid = (unsigned)(n_tiles < ctas_per_device ? n_tiles : ctas_per_device);
    void* args[] = {(void*)&P};
    OPTK_CUDA(cudaLaunchKernel(kernel, dim3(grid), dim3(THREADS), args, smem, stream));
    return OPTK_OK;
}

// (threads per CTA, CTAs per SM) of the pipeline; OPTK_TMA_SHAPE selects one of the measured
// alternatives
int tma_shape() {
    static const int shape = [] {
        const char* e = getenv("OPTK_TMA_SHAPE");
        const int v = e ? atoi(e) : 3;
        return (v >= 0 && v <= 3) ? v : 3;
    }();
    return shape;
}

}  // namespace

int tma_tile_rays() {
    static const int tiles[4] = {512, 384, 256, 256};
    return tiles[tma_shape()];
}

int launch_trace_tma(const TraceParams& P, cudaStream_t stream) {
    switch (tma_shape()) {
        case 1: return launch_shape<192, 3>(P, stream);
        case 2: return launch_shape<128, 5>(P, stream);
        case 3: return launch_shape<128, 4>(P, stream);
        default: return launch_shape<256, 2>(P, stream);
    }
}

}  // namespace optk
