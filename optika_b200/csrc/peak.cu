// FP64 DFMA peak micro-benchmark: the roofline denominator MEASURED_PEAKS.json
// does not record.  148 SMs x resident CTAs of independent DFMA chains.
#include "common.cuh"
#include "params.cuh"

namespace optk {

__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iterations) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0;
    double a4 = a0 + 4.0, a5 = a0 + 5.0, a6 = a0 + 6.0, a7 = a0 + 7.0;
    const double b = 1.0000000001, c = 1e-12;
    for (int i = 0; i < iterations; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
            a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
        }
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 12345.678) out[0] = s;  // never true; keeps the chains alive
}

int measure_fp64_peak(double* flops, cudaStream_t stream) {
    int device = 0, sms = 0;
    OPTK_CUDA(cudaGetDevice(&device));
    OPTK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    double* out = nullptr;
    OPTK_CUDA(cudaMalloc((void**)&out, 8));
    cudaEvent_t e0, e1;
    OPTK_CUDA(cudaEventCreate(&e0));
    OPTK_CUDA(cudaEventCreate(&e1));
    const int block = 256, grid = sms * 8, iterations = 4096;
    dfma_kernel<<<grid, block, 0, stream>>>(out, 64);  // warm up
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        OPTK_CUDA(cudaEventRecord(e0, stream));
        dfma_kernel<<<grid, block, 0, stream>>>(out, iterations);
        OPTK_CUDA(cudaEventRecord(e1, stream));
        OPTK_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        OPTK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = 2.0 * 64.0 * iterations * (double)block * grid;  // 64 FMAs per iteration per thread
        const double rate = fl / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    OPTK_CUDA(cudaGetLastError());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *flops = best;
    return OPTK_OK;
}

// ---------------------------------------------------------------------------
// Bandwidth of the trace kernel's ACCESS PATTERN with no arithmetic: ten fp64 arrays
// and one byte mask in, the same out, two rays per thread with 128-bit accesses.
// Separates "the memory system cannot stream 22 interleaved arrays faster" from
// "the kernel does not issue fast enough".
// ---------------------------------------------------------------------------
struct SoaPtrs {
    const double* in[10];
    double* out[10];
    const unsigned char* mask_in;
    unsigned char* mask_out;
};

__global__ void __launch_bounds__(256) soa_copy_kernel(const __grid_constant__ SoaPtrs p, long long n) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i + 1 >= n) return;
    double2 v[10];
#pragma unroll
    for (int f = 0; f < 10; ++f) v[f] = __ldg(reinterpret_cast<const double2*>(p.in[f] + i));
    const uchar2 m = *reinterpret_cast<const uchar2*>(p.mask_in + i);
#pragma unroll
    for (int f = 0; f < 10; ++f) *reinterpret_cast<double2*>(p.out[f] + i) = v[f];
    *reinterpret_cast<uchar2*>(p.mask_out + i) = m;
}

int measure_soa_copy(long long n_rays, double* gbytes_per_second, cudaStream_t stream) {
    n_rays &= ~1LL;
    SoaPtrs p;
    char* base = nullptr;
    const size_t field_bytes = (size_t)n_rays * 8, mask_bytes = ((size_t)n_rays + 255) & ~(size_t)255;
    const size_t total = 20 * field_bytes + 2 * mask_bytes;
    cudaError_t e = cudaMalloc((void**)&base, total);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes) failed: %s", total, cudaGetErrorString(e));
        return OPTK_ERR_NOMEM;
    }
    OPTK_CUDA(cudaMemsetAsync(base, 0, total, stream));
    for (int f = 0; f < 10; ++f) {
        p.in[f] = (const double*)(base + f * field_bytes);
        p.out[f] = (double*)(base + (10 + f) * field_bytes);
    }
    p.mask_in = (const unsigned char*)(base + 20 * field_bytes);
    p.mask_out = (unsigned char*)(base + 20 * field_bytes + mask_bytes);
    cudaEvent_t e0, e1;
    OPTK_CUDA(cudaEventCreate(&e0));
    OPTK_CUDA(cudaEventCreate(&e1));
    const int block = 256;
    const long long grid = (n_rays / 2 + block - 1) / block;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        OPTK_CUDA(cudaEventRecord(e0, stream));
        soa_copy_kernel<<<(unsigned)grid, block, 0, stream>>>(p, n_rays);
        OPTK_CUDA(cudaEventRecord(e1, stream));
        OPTK_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        OPTK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double rate = 162.0 * (double)n_rays / (ms * 1e-3) / 1e9;
        if (rep > 0 && rate > best) best = rate;
    }
    OPTK_CUDA(cudaGetLastError());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(base);
    *gbytes_per_second = best;
    return OPTK_OK;
}

}  // namespace optk
