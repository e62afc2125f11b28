// Streamlined kernels for long or expensive surface lists (six surfaces, toroids, ...), whose
// walk is FP64-latency bound rather than issue bound: compiled for two resident CTAs per SM,
// i.e. up to 128 registers and no spills (16 warps instead of 24).  Only the generate + trace +
// bin kernels gain from it: cfg-1 image 8.79 -> 7.90 ms per 1e8 rays, cfg 5 733 -> 710 ms; the
// same walk fed from broadcast grids (5.45 -> 6.58 ms) or writing rays (5.14 -> 5.82 ms) loses,
// and so does cfg 2 (three cheap surfaces: 16.7 -> 19.4 ms per 4e8 rays).  launch_trace picks.
#include "trace_impl.cuh"

namespace optk {

trace_kernel_t select_heavy_kernel(bool grid, bool acc, bool image) {
#define OPTK_PICK(G, A, I) \
    if (grid == G && acc == A && image == I) return (trace_kernel_t)trace_kernel<2, 2, true, false, false, A, I, G ? 1 : 0>;
    OPTK_PICK(true, false, true)
    OPTK_PICK(true, true, true)
#undef OPTK_PICK
    return nullptr;
}

}  // namespace optk
