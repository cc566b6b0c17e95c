// FmIndex<u64, BlockN<u64>, *> for N = 2..6: kernel instantiations and their launchers (engine.cuh).
#include "engine.cuh"

namespace svfm {
SVFM_DEFINE_TYPE_OPS(ops_p64_v64, uint64_t, 64)
}  // namespace svfm
