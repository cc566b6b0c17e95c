// FmIndex<u64, BlockN<u128>, *> for N = 2..6: kernel instantiations and their launchers (engine.cuh).
#include "engine.cuh"

namespace svfm {
SVFM_DEFINE_TYPE_OPS(ops_p64_v128, uint64_t, 128)
}  // namespace svfm
