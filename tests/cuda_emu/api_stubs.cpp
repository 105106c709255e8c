// api_stubs.cpp -- TEST ONLY, part of libbonejthickness_emu.so (the whole C-ABI on the CPU emulation of CUDA):
// the dynamic shared memory of edt_tile_kernel.
#include "common.cuh"

namespace bjt {
alignas(16) uint32_t sm_raw[(128 * 1024) / 4];
namespace disc_detail { alignas(16) uint32_t zf_smem[(224 * 1024) / 4], dk_smem[(128 * 1024) / 4]; }   // paint_disc.cu
}  // namespace bjt
