// FmIndex<u32, BlockN<u32>, *> for N = 2..6: kernel instantiations and their launchers (engine.cuh).
#include "engine.cuh"

namespace svfm {
SVFM_DEFINE_TYPE_OPS(ops_p32_v32, uint32_t, 32)
}  // namespace svfm
