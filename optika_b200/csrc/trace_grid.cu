// Kernel 1 with the on-device ray generator in front (optk_trace_grid): the rays of a
// separable, optionally jittered, vertex grid are created in registers, traced and
// (optionally) binned without a single input byte read from HBM.
#include "trace_impl.cuh"

namespace optk {

trace_kernel_t select_grid_kernel(bool full, bool acc, bool image) {
#define OPTK_PICK(A, I)                                                                             \
    if (acc == A && image == I)                                                                     \
        return full ? (trace_kernel_t)trace_kernel<OPTK_FULL_MINB, 2, true, false, false, A, I, true>            \
                    : (trace_kernel_t)trace_kernel<4, 1, false, false, false, A, I, true>;
    OPTK_PICK(false, false)
    OPTK_PICK(true, false)
    OPTK_PICK(false, true)
    OPTK_PICK(true, true)
#undef OPTK_PICK
    return nullptr;
}

}  // namespace optk
