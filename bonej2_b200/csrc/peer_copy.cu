// peer_copy.cu -- strided device-to-device copies done by the SMs, local or out of a peer GPU's memory over NVLink.
//
// The z-slab split needs two layout changes per run (the all-to-all transpose in front of the EDT z pass,
// and the transpose back together with the max-radius halo).  Each is, per peer, one strided block: `height` rows of
// `width` bytes.  One launch moves the blocks of ALL peers (grid.y = block): 16-byte loads out of the mapped peer
// buffers (CUDA IPC between processes, plain peer access inside one), eight in flight per thread, written to local
// memory.  Done by SMs rather than by the copy engines so that the pulls of a run do not queue behind the download of
// the previous map on the same engines, and so that seven peers are read at once.
#include "common.cuh"

namespace bjt {

namespace {

constexpr int SEG16 = 4096;                 // 16-byte units per segment (64 KB): one block copies a segment at a time
constexpr int CP_THREADS = 256;
constexpr int CP_UNROLL = 8;

__global__ void __launch_bounds__(CP_THREADS) copy2d_batch_kernel(CopyBatch b) {
    const CopyDesc d = b.d[blockIdx.y];
    const unsigned cpr = (unsigned)((d.width16 + SEG16 - 1) / SEG16);          // segments per row
    const unsigned long long nseg = (unsigned long long)d.height * cpr;
    for (unsigned long long seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
        const unsigned long long row = seg / cpr;
        const size_t c0 = (size_t)(seg % cpr) * SEG16;
        const size_t n = min((size_t)SEG16, d.width16 - c0);
        const uint4 *s = d.src + row * d.spitch16 + c0;
        uint4 *q = d.dst + row * d.dpitch16 + c0;
        for (size_t i0 = threadIdx.x; i0 < n; i0 += (size_t)CP_THREADS * CP_UNROLL) {
            uint4 v[CP_UNROLL];
#pragma unroll
            for (int u = 0; u < CP_UNROLL; ++u) {
                const size_t i = i0 + (size_t)u * CP_THREADS;
                if (i < n) v[u] = s[i];
            }
#pragma unroll
            for (int u = 0; u < CP_UNROLL; ++u) {
                const size_t i = i0 + (size_t)u * CP_THREADS;
                if (i < n) q[i] = v[u];
            }
        }
    }
}

}  // namespace

// every block of the batch must be 16-byte aligned in pointers, pitches and width (the callers check)
int launch_copy2d_batch(const CopyBatch &b, int sm_count, cudaStream_t s) {
    if (b.n <= 0) return 0;
    const unsigned per = (unsigned)max(8, (sm_count * 8 + b.n - 1) / b.n);     // ~8 blocks per SM over the whole batch
    copy2d_batch_kernel<<<dim3(per, (unsigned)b.n), CP_THREADS, 0, s>>>(b);
    return 1;
}

}  // namespace bjt
