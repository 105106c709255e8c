// api_stubs.cpp -- TEST ONLY, part of libbonejthickness_emu.so (the whole C-ABI on the CPU emulation of CUDA).
// The dynamic shared memory of edt_tile_kernel, and loud stand-ins for the two kernel files the emulation does not
// cover (they use __activemask / __match_any_sync in divergent code, which a fiber model cannot reproduce).
#include "common.cuh"

namespace bjt {
alignas(16) uint32_t sm_raw[(128 * 1024) / 4];

static int not_emulated(const char *what) {
    fprintf(stderr, "libbonejthickness_emu: %s is not covered by the CPU emulation\n", what);
    abort();
    return 0;
}
int launch_label_stats(const float *, const int32_t *, size_t, int, const double *, double *, uint32_t *, cudaStream_t) { return not_emulated("segstats.cu"); }
int launch_slice_stats(const float *, int, int, int, int, int, int, int, double *, cudaStream_t) { return not_emulated("segstats.cu"); }
int launch_pad_binary(const uint8_t *, int, int, int, uint8_t *, cudaStream_t) { return not_emulated("ridgepoints.cu"); }
int launch_ridge_points_map(const uint32_t *, int, int, int, float *, float *, float *, uint32_t *, cudaStream_t) { return not_emulated("ridgepoints.cu"); }
int launch_ridge_points_collect(const float *, size_t, float, unsigned long long *, unsigned long long, unsigned long long *, cudaStream_t) { return not_emulated("ridgepoints.cu"); }
}  // namespace bjt
