// Kernel 1 with the on-device ray generator in front (optk_trace_grid): the rays of a
// separable, optionally jittered, vertex grid are created in registers, traced and
// (optionally) binned without a single input byte read from HBM.
#include "trace_impl.cuh"

namespace optk {

// curvilinear grids (2-D vertex arrays) have their own instantiations: the out-of-line bilinear
// sampler's call sites cost the separable kernels 5 % even when never taken
trace_kernel_t select_grid_kernel(bool full, bool acc, bool image, bool curvilinear) {
#define OPTK_PICK(A, I, G)                                                                          \
    if (acc == A && image == I && curvilinear == (G == 2))                                          \
        return full ? (trace_kernel_t)trace_kernel<OPTK_FULL_MINB, 2, true, false, false, A, I, G>  \
                    : (trace_kernel_t)trace_kernel<4, 1, false, false, false, A, I, G>;
    OPTK_PICK(false, false, 1)
    OPTK_PICK(true, false, 1)
    OPTK_PICK(false, true, 1)
    OPTK_PICK(true, true, 1)
    OPTK_PICK(false, false, 2)
    OPTK_PICK(true, false, 2)
    OPTK_PICK(false, true, 2)
    OPTK_PICK(true, true, 2)
#undef OPTK_PICK
    return nullptr;
}

}  // namespace optk
