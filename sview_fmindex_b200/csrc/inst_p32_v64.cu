// FmIndex<u32, BlockN<u64>, *> for N = 2..6: kernel instantiations and their launchers (engine.cuh).
#include "engine.cuh"

namespace svfm {
SVFM_DEFINE_TYPE_OPS(ops_p32_v64, uint32_t, 64)
}  // namespace svfm
