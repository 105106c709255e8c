// paint_scatter.cu -- brute-force largest-inscribed-sphere painting (path 2), sm_100a.
//
// The literal form of sc.fiji.localThickness.Local_Thickness_Parallel: every ridge
// point (i,j,k,R) raises all voxels with (i1-i)^2+(j1-j)^2+(k1-k)^2 <= R to at least R.  The Java
// code guards the max with a per-slice lock; here it is an integer atomicMax, so the result is
// independent of the order in which points are painted.  Cost is sum over ridge points of the ball
// volume, so this path is the always-correct fallback of the separable painter (paint_sep.cu) and
// the reference point it is tested against; it is not the path used on large volumes.
#include "common.cuh"

namespace bjt {

__global__ void compact_ridge_kernel(const uint32_t *__restrict__ ridge, int w, int h, int d, uint4 *__restrict__ points,
                                     unsigned long long *__restrict__ count) {
    const size_t n = (size_t)w * h * d;
    const int lane = threadIdx.x & 31;
    for (size_t base = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) - lane; base < n;
         base += (size_t)gridDim.x * blockDim.x) {
        const size_t i = base + lane;
        const uint32_t v = i < n ? ridge[i] : 0u;
        const unsigned ballot = __ballot_sync(0xFFFFFFFFu, v != 0u);
        if (!ballot) continue;
        unsigned long long off = 0;
        if (lane == 0) off = atomicAdd(count, (unsigned long long)__popc(ballot));
        off = __shfl_sync(0xFFFFFFFFu, off, 0);
        if (v && points) {
            const size_t wh = (size_t)w * h;
            const uint32_t z = (uint32_t)(i / wh);
            const uint32_t rem = (uint32_t)(i % wh);
            const uint32_t y = rem / (uint32_t)w, x = rem % (uint32_t)w;
            points[off + __popc(ballot & ((1u << lane) - 1u))] = make_uint4(x | (y << 16), z, v, 0u);
        }
    }
}

int launch_compact_ridge(const uint32_t *ridge, int w, int h, int d, uint4 *points, unsigned long long *count,
                         cudaStream_t s) {
    const size_t n = (size_t)w * h * d;
    const unsigned grid = (unsigned)min((size_t)148 * 16, (n + 255) / 256);
    compact_ridge_kernel<<<grid ? grid : 1, 256, 0, s>>>(ridge, w, h, d, points, count);
    return 1;
}

// one warp per ridge point; lanes sweep x
__global__ void __launch_bounds__(256) paint_scatter_kernel(const uint4 *__restrict__ points, unsigned long long npoints,
                                                             int w, int h, int d, uint32_t *__restrict__ rsq) {
    const int lane = threadIdx.x & 31;
    const size_t wh = (size_t)w * h;
    for (unsigned long long p = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); p < npoints;
         p += (unsigned long long)gridDim.x * (blockDim.x >> 5)) {
        const uint4 pt = __ldg(points + p);
        const int i = (int)(pt.x & 0xFFFFu), j = (int)(pt.x >> 16), k = (int)pt.y;
        const uint32_t R = pt.z;
        const int r = (int)isqrt_ceil(R);          // rInt = ceil(r)
        const int k0 = max(k - r, 0), k1 = min(k + r, d - 1);
        const int j0 = max(j - r, 0), j1 = min(j + r, h - 1);
        for (int kk = k0; kk <= k1; ++kk) {
            const uint32_t dk2 = (uint32_t)((kk - k) * (kk - k));
            for (int jj = j0; jj <= j1; ++jj) {
                const uint32_t djk2 = dk2 + (uint32_t)((jj - j) * (jj - j));
                if (djk2 > R) continue;
                const int hw = (int)isqrt_floor(R - djk2);
                const int i0 = max(i - hw, 0), i1 = min(i + hw, w - 1);
                uint32_t *row = rsq + kk * wh + (size_t)jj * w;
                for (int ii = i0 + lane; ii <= i1; ii += 32)
                    if (row[ii] < R) atomicMax(row + ii, R);
            }
        }
    }
}

int launch_paint_scatter(const uint4 *points, unsigned long long npoints, int w, int h, int d, uint32_t *rsq,
                         cudaStream_t s) {
    if (!npoints) return 0;
    const unsigned long long blocks = (npoints + 7) / 8;
    const unsigned grid = (unsigned)min(blocks, (unsigned long long)148 * 64);
    paint_scatter_kernel<<<grid, 256, 0, s>>>(points, npoints, w, h, d, rsq);
    return 1;
}

__global__ void fill_u32_kernel(uint32_t *p, size_t n, uint32_t v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
int launch_fill_u32(uint32_t *p, size_t n, uint32_t v, cudaStream_t s) {
    const unsigned grid = (unsigned)min((size_t)148 * 16, (n + 255) / 256);
    fill_u32_kernel<<<grid ? grid : 1, 256, 0, s>>>(p, n, v);
    return 1;
}

}  // namespace bjt
