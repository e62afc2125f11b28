// Detector physics after binning: the Monte-Carlo electron kernel.
//
// Restates optika/sensors/materials/_ramanathan_2020/_ramanathan_2020.py:762-876
// (_electrons_measured_numba, a numba prange over image planes): for every photon absorbed in a pixel,
//   1. the number of electron-hole pairs -- drawn from the tabulated pair-number distribution below
//      50 eV (:826-833), from a rounded normal with Fano variance above (:835-840);
//   2. the absorption depth, exponential truncated to the substrate (:842-848);
//   3. the charge-collection efficiency at that depth: linear ramp over the implant (:850-853), applied
//      as a binomial thinning of the pairs (:855);
//   4. for every surviving electron, a Gaussian lateral step of width z_ff sqrt(1 - z / z_ff) from the
//      photon's position inside the pixel, rounded to whole pixels (:857-872), deposited with toroidal
//      wrap or dropped at the edge (:874-880).
// One thread per pixel walks that pixel's photons; electrons that stay in the pixel accumulate in a
// register, the others are global integer atomics on the neighbours.  Random numbers are counter
// based (Philox4x32-10, key = seed, counter = {pixel, photon, draw}): the result does not depend on
// the launch geometry, and the NumPy oracle (oracle/detector.py) reproduces it count for count.
#include "common.cuh"
#include "trace_impl.cuh"  // philox4x32_10

namespace optk {

namespace {

__device__ __forceinline__ double uniform53(uint32_t hi, uint32_t lo) {
    const unsigned long long bits = ((unsigned long long)hi << 21) | (unsigned long long)(lo >> 11);
    return ((double)bits + 0.5) * 1.1102230246251565e-16;  // 2^-53: strictly inside (0, 1)
}

__device__ __forceinline__ void words(unsigned long long pixel, uint32_t photon, uint32_t draw, unsigned long long seed,
                                      uint32_t (&x)[4]) {
    philox4x32_10((uint32_t)pixel, (uint32_t)(pixel >> 32), photon, draw, (uint32_t)seed, (uint32_t)(seed >> 32), x);
}

// Box-Muller on two uniforms in (0, 1)
__device__ __forceinline__ void normal_pair(double u1, double u2, double& z0, double& z1) {
    const double r = sqrt(-2.0 * log(u1));
    double s, c;
    sincos(6.283185307179586 * u2, &s, &c);
    z0 = r * c;
    z1 = r * s;
}

__device__ __forceinline__ int wrap_index(long long i, int n) {
    long long m = i % n;
    return (int)(m < 0 ? m + n : m);
}

__global__ void __launch_bounds__(128)
electrons_kernel(int n_plane, int n_x, int n_y, const optk_ccd_plane_t* __restrict__ planes,
                 const long long* __restrict__ photons, unsigned long long* __restrict__ electrons, int wrap,
                 unsigned long long seed) {
    const long long n_pixel = (long long)n_x * n_y;
    const long long total = n_pixel * n_plane;
    for (long long pixel = (long long)blockIdx.x * blockDim.x + threadIdx.x; pixel < total;
         pixel += (long long)gridDim.x * blockDim.x) {
        const long long num_photon = photons[pixel];
        if (num_photon <= 0) continue;
        const int i = (int)(pixel / n_pixel);
        const long long in_plane = pixel - (long long)i * n_pixel;
        const int x = (int)(in_plane / n_y), y = (int)(in_plane - (long long)x * n_y);
        const optk_ccd_plane_t P = planes[i];
        const double a = P.absorption, W = P.thickness_implant, h_0 = P.cce_backsurface;
        const double z_substrate = P.thickness_substrate, z_ff = z_substrate - P.thickness_depletion;
        const double d = a > 0 ? 1.0 / a : 0.0;
        const double fraction_absorbed = 1.0 - exp(-a * z_substrate);
        const double mean_inf = P.energy / P.energy_pair_inf, std_inf = sqrt(P.fano_inf * mean_inf);
        const bool low_energy = P.energy <= 50.0;
        const bool can_diffuse = P.width_pixel_x > 0 && P.width_pixel_y > 0;
        unsigned long long* plane_out = electrons + (long long)i * n_pixel;
        unsigned long long here = 0;
        for (long long j = 0; j < num_photon; ++j) {
            uint32_t w0[4], w1[4];
            words((unsigned long long)pixel, (uint32_t)j, 0u, seed, w0);
            words((unsigned long long)pixel, (uint32_t)j, 1u, seed, w1);
            long long n_ij;
            if (low_energy) {
                const double x_ij = uniform53(w0[0], w0[1]);
                int k = 0;
                while (k < P.n_pmf - 1 && !(P.cmf[k] > x_ij)) ++k;  // first k with cmf[k] > x, else the last
                n_ij = (long long)P.n_values[k];
            } else {
                double z0, z1;
                normal_pair(uniform53(w0[0], w0[1]), uniform53(w0[2], w0[3]), z0, z1);
                n_ij = (long long)rint(mean_inf + std_inf * z0);
            }
            if (n_ij <= 0) continue;
            const double y_ij = uniform53(w1[0], w1[1]);
            const double z_ij = a > 0 ? -d * log(1.0 - y_ij * fraction_absorbed) : y_ij * z_substrate;
            const double h_ij = (z_ij < W) ? (W > 0 ? h_0 + (1.0 - h_0) * z_ij / W : 1.0) : 1.0;
            const double u = ((double)w1[2] + 0.5) * 2.3283064365386963e-10 - 0.5;  // uniform(-0.5, 0.5): 2^-32
            const double v = ((double)w1[3] + 0.5) * 2.3283064365386963e-10 - 0.5;
            const bool diffuses = z_ij < z_ff && can_diffuse;
            const double w = diffuses ? z_ff * sqrt(fmax(1.0 - z_ij / z_ff, 0.0)) : 0.0;
            const double sx = diffuses ? w / P.width_pixel_x : 0.0, sy = diffuses ? w / P.width_pixel_y : 0.0;
            for (long long e = 0; e < n_ij; ++e) {
                if (h_ij < 1.0) {  // binomial(n, h) as n Bernoulli trials
                    uint32_t ws[4];
                    words((unsigned long long)pixel, (uint32_t)j, 0x80000000u + 2u + (uint32_t)e, seed, ws);
                    if (!(uniform53(ws[0], ws[1]) < h_ij)) continue;
                }
                if (!diffuses) {
                    ++here;
                    continue;
                }
                uint32_t we[4];
                words((unsigned long long)pixel, (uint32_t)j, 2u + (uint32_t)e, seed, we);
                double zp, zq;
                normal_pair(uniform53(we[0], we[1]), uniform53(we[2], we[3]), zp, zq);
                const long long p = (long long)rint(u + sx * zp), q = (long long)rint(v + sy * zq);
                if (p == 0 && q == 0) {
                    ++here;
                    continue;
                }
                const long long x_e = x + p, y_e = y + q;
                if (wrap) {
                    atomicAdd(plane_out + (long long)wrap_index(x_e, n_x) * n_y + wrap_index(y_e, n_y), 1ULL);
                } else if (x_e >= 0 && x_e < n_x && y_e >= 0 && y_e < n_y) {
                    atomicAdd(plane_out + x_e * n_y + y_e, 1ULL);
                }  // otherwise the electron diffused off the sensor and is lost
            }
        }
        if (here) atomicAdd(plane_out + in_plane, here);
    }
}

}  // namespace

int launch_electrons(int n_plane, int n_x, int n_y, const optk_ccd_plane_t* planes_device, const long long* photons,
                     unsigned long long* electrons, int wrap, unsigned long long seed, cudaStream_t stream) {
    const long long total = (long long)n_plane * n_x * n_y;
    if (total <= 0) return OPTK_OK;
    int device = 0, sms = 0;
    OPTK_CUDA(cudaGetDevice(&device));
    OPTK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int block = 128;
    long long grid = (total + block - 1) / block;
    const long long resident = (long long)sms * 16;  // grid-stride beyond 16 CTAs per SM
    if (grid > resident) grid = resident;
    electrons_kernel<<<(unsigned)grid, block, 0, stream>>>(n_plane, n_x, n_y, planes_device, photons, electrons, wrap, seed);
    OPTK_CUDA(cudaGetLastError());
    return OPTK_OK;
}

}  // namespace optk
