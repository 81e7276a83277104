// The "lean general" megakernel (kFeatLean): its own translation unit so that it compiles in parallel with the others.
#define RT_RECT_OUTLINE 1  // path_logic.cuh: this kernel is bound by instruction fetch
#include "../abi/kernel_table.hpp"
#include "../device/render_kernel.cuh"

namespace rtiow {
namespace {
template <bool F, bool Q>
KernelVariant lean_by_threads(uint32_t threads) {
    switch (threads) {
        case 512: return {render_kernel<true, F, Q, kFeatLean, 512, 1>, 512};
        case 768: return {render_kernel<true, F, Q, kFeatLean, 768, 1>, 768};
        default: return {nullptr, 0};
    }
}
}  // namespace

KernelVariant pick_lean_smem(bool frames, bool fast, uint32_t threads) {
    if (frames) return fast ? lean_by_threads<true, true>(threads) : lean_by_threads<true, false>(threads);
    return fast ? lean_by_threads<false, true>(threads) : lean_by_threads<false, false>(threads);
}
}  // namespace rtiow
