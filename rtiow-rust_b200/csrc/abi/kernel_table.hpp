// The megakernel instantiations live in their own translation units (csrc/kernels/*.cu) so that
// they compile in parallel; the ABI implementation picks one through these functions.
#pragma once
#include <cstdint>

#include "../device/path_logic.cuh"

namespace rtiow {

typedef void (*kernel_fn)(const KParams);

struct KernelVariant {
    kernel_fn fn;   // nullptr: no such instantiation
    int threads;
};

// render_kernel (render_kernel.cuh): one path per lane.  threads in {256, 512, 768}.
// profile: 0 general, 1 spheres-only (path_logic.cuh kFeatSpheres), 2 rect lists (kFeatRects); threads 1024 for 1 and 2 only
KernelVariant pick_plain_smem(bool frames, bool fast, uint32_t profile, uint32_t threads);
KernelVariant pick_plain_global(bool frames, bool fast, uint32_t profile, uint32_t threads);
// the general kernel without the features the reference's own scenes never use (path_logic.cuh kFeatLean); shared-memory
// scenes only, threads in {512, 768}
KernelVariant pick_lean_smem(bool frames, bool fast, uint32_t threads);

}  // namespace rtiow
