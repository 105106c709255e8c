// paint_sep.cu -- separable largest-inscribed-sphere painter (placeholder until the real one lands)
#include "common.cuh"
namespace bjt {
size_t paint_sep_workspace_bytes(int, int, int, int) { return 0; }
int launch_paint_sep(const uint32_t *, int, int, int, int, void *, uint32_t *, uint32_t *, cudaStream_t) { return 0; }
}
