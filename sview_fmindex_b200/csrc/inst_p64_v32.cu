// FmIndex<u64, BlockN<u32>, *> for N = 2..6: kernel instantiations and their launchers (engine.cuh).
#include "engine.cuh"

namespace svfm {
SVFM_DEFINE_TYPE_OPS(ops_p64_v32, uint64_t, 32)
}  // namespace svfm
