// Included by plain_smem.cu / plain_global.cu with RTIOW_PLAIN_SMEM and RTIOW_PLAIN_NAME defined.
#include "../abi/kernel_table.hpp"
#include "../device/render_kernel.cuh"

namespace rtiow {
namespace {
template <bool F, bool Q, bool L>
KernelVariant by_threads(uint32_t threads) {
    switch (threads) {
        case 256: return {render_kernel<RTIOW_PLAIN_SMEM, F, Q, L, 256, 1>, 256};
        case 512: return {render_kernel<RTIOW_PLAIN_SMEM, F, Q, L, 512, 1>, 512};
        case 768: return {render_kernel<RTIOW_PLAIN_SMEM, F, Q, L, 768, 1>, 768};
        case 1024: if (L) return {render_kernel<RTIOW_PLAIN_SMEM, F, Q, true, 1024, 1>, 1024}; return {nullptr, 0};
        default: return {nullptr, 0};
    }
}
}  // namespace

// lean: the spheres-only specialisation (path_logic.cuh SceneT); never combined with frames
KernelVariant RTIOW_PLAIN_NAME(bool frames, bool fast, bool lean, uint32_t threads) {
    if (lean && !frames) return fast ? by_threads<false, true, true>(threads) : by_threads<false, false, true>(threads);
    if (frames) return fast ? by_threads<true, true, false>(threads) : by_threads<true, false, false>(threads);
    return fast ? by_threads<false, true, false>(threads) : by_threads<false, false, false>(threads);
}
}  // namespace rtiow
