// Streamlined (two rays per thread) kernels for systems with non-unit efficiencies: measured
// mirrors / rulings and groove profiles (optika/surfaces.py:175-179).  Same code as the other
// streamlined kernels plus one out-of-line call per ray at the surfaces that need it; kept
// in their own instantiations because the mere presence of the call site costs the
// unit-efficiency kernels 2-3 %.
#include "trace_impl.cuh"

namespace optk {

trace_kernel_t select_efficiency_kernel(bool grid, bool dense, bool acc, bool image) {
#define OPTK_PICK(G, D, A, I)                                    \
    if (grid == G && dense == D && acc == A && image == I)       \
        return (trace_kernel_t)trace_kernel<OPTK_FULL_MINB, 2, true, D, false, A, I, G ? 1 : 0, true>;
    OPTK_PICK(false, false, false, false)
    OPTK_PICK(false, false, true, false)
    OPTK_PICK(false, false, false, true)
    OPTK_PICK(false, false, true, true)
    OPTK_PICK(false, true, false, false)
    OPTK_PICK(false, true, true, false)
    OPTK_PICK(false, true, false, true)
    OPTK_PICK(false, true, true, true)
    OPTK_PICK(true, false, false, false)
    OPTK_PICK(true, false, true, false)
    OPTK_PICK(true, false, false, true)
    OPTK_PICK(true, false, true, true)
#undef OPTK_PICK
    return nullptr;
}

}  // namespace optk
