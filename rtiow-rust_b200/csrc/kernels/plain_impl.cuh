// Included by plain_smem.cu / plain_global.cu with RTIOW_PLAIN_SMEM and RTIOW_PLAIN_NAME defined.
#include "../abi/kernel_table.hpp"
#include "../device/render_kernel.cuh"

namespace rtiow {
namespace {
template <bool F, bool Q, uint32_t M>
KernelVariant by_threads(uint32_t threads) {
    switch (threads) {
        case 256: return {render_kernel<RTIOW_PLAIN_SMEM, F, Q, M, 256, 1>, 256};
        case 512: return {render_kernel<RTIOW_PLAIN_SMEM, F, Q, M, 512, 1>, 512};
        case 768: return {render_kernel<RTIOW_PLAIN_SMEM, F, Q, M, 768, 1>, 768};
        case 1024:  // 64 registers: only the specialised kernels fit
            if (M != SF_ALL) return {render_kernel<RTIOW_PLAIN_SMEM, F, Q, M == SF_ALL ? kFeatSpheres : M, 1024, 1>, 1024};
            return {nullptr, 0};
        default: return {nullptr, 0};
    }
}
}  // namespace

// profile: 0 = any scene, 1 = spheres-only scenes (kFeatSpheres), 2 = rect-list scenes (kFeatRects, no accel, so
// `fast` does not apply); 1 and 2 are never combined with frames
KernelVariant RTIOW_PLAIN_NAME(bool frames, bool fast, uint32_t profile, uint32_t threads) {
    if (!frames && profile == 1u) return fast ? by_threads<false, true, kFeatSpheres>(threads) : by_threads<false, false, kFeatSpheres>(threads);
    if (!frames && profile == 2u) return by_threads<false, false, kFeatRects>(threads);
    if (frames) return fast ? by_threads<true, true, SF_ALL>(threads) : by_threads<true, false, SF_ALL>(threads);
    return fast ? by_threads<false, true, SF_ALL>(threads) : by_threads<false, false, SF_ALL>(threads);
}
}  // namespace rtiow
