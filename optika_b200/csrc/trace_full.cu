// Instantiations of the streamlined (FULL operator, two rays per thread) trace kernels.
#include "trace_impl.cuh"

namespace optk {

trace_kernel_t select_full_kernel(bool dense, bool vec, bool acc, bool image) {
#define OPTK_PICK(D, V, A, I) \
    if (dense == D && vec == V && acc == A && image == I) return (trace_kernel_t)trace_kernel<OPTK_FULL_MINB, 2, true, D, V, A, I>;
    OPTK_PICK(true, true, false, false)
    OPTK_PICK(true, true, true, false)
    OPTK_PICK(true, true, false, true)
    OPTK_PICK(true, true, true, true)
    OPTK_PICK(true, false, false, false)
    OPTK_PICK(true, false, true, false)
    OPTK_PICK(true, false, false, true)
    OPTK_PICK(true, false, true, true)
    OPTK_PICK(false, false, false, false)
    OPTK_PICK(false, false, true, false)
    OPTK_PICK(false, false, false, true)
    OPTK_PICK(false, false, true, true)
#undef OPTK_PICK
    return nullptr;
}

}  // namespace optk
