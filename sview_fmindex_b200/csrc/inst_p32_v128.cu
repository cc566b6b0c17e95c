// FmIndex<u32, BlockN<u128>, *> for N = 2..6: kernel instantiations and their launchers (engine.cuh).
#include "engine.cuh"

namespace svfm {
SVFM_DEFINE_TYPE_OPS(ops_p32_v128, uint32_t, 128)
}  // namespace svfm
