// finish.cu -- everything after painting, sm_100a:
//   (float)(2*sqrt(R))                      tail of Local_Thickness_Parallel            
//   Clean_Up_Local_Thickness                setFlag / averageInteriorNeighbors          (A.4)
//   MaskThicknessMapWithOriginal.trimOverhang                                            (A.5)
//   0 -> NaN, x (float)pixelWidth           LocalThicknessWrapper calibration tail       (A.6)
//   ij.process.StackStatistics              count / sum / sum^2 / min / max in double    (A.7)
// as driven by ThicknessWrapper.createMap / addMapResults (ThicknessWrapper.java:158-196).
//
// One fused kernel.  A block owns 32 x 8 x 8 output voxels and stages the painted r^2 of the
// (36 x 12 x 12) tile with its 2-voxel halo in shared memory (out-of-image cells hold a marker that
// counts as "not zero", as the reference's look() returns -1 there).  From the tile it derives the
// "interior" flag (non-zero, no zero 26-neighbour) on the 1-voxel halo, then the value of each output
// voxel: interior voxels keep 2*sqrt(R); surface voxels take the float32 mean of their interior
// neighbours visited in the order faces, edges, corners (the order fixes the last ulp); then mask,
// NaN, calibration, and per-block partials of the statistics.  HBM sees r^2 once (plus halo re-reads
// that hit L2), the original once and the map once: 9 B/voxel of actual traffic.
#include "common.cuh"
#include <math.h>

namespace bjt {

constexpr int FTX = 32, FTY = 8, FTZ = 8;
constexpr int HX = FTX + 4, HY = FTY + 4, HZ = FTZ + 4;       // r^2 tile with 2-voxel halo
constexpr int IX = FTX + 2, IY = FTY + 2, IZ = FTZ + 2;       // interior flags with 1-voxel halo
constexpr uint32_t OUTSIDE = 0xFFFFFFFFu;

// neighbour visiting order of the cleanup average: 6 faces, 12 edges, 8 corners (dx, dy, dz)
__constant__ int8_t NB26[26][3] = {
    {0, 0, -1}, {0, 0, 1}, {0, -1, 0}, {0, 1, 0}, {-1, 0, 0}, {1, 0, 0},
    {0, 1, -1}, {0, 1, 1}, {1, -1, 0}, {1, 1, 0}, {-1, 0, 1}, {1, 0, 1},
    {0, -1, -1}, {0, -1, 1}, {-1, -1, 0}, {-1, 1, 0}, {-1, 0, -1}, {1, 0, -1},
    {1, 1, 1}, {1, -1, 1}, {-1, 1, 1}, {-1, -1, 1},
    {1, 1, -1}, {1, -1, -1}, {-1, 1, -1}, {-1, -1, -1}};

__device__ __forceinline__ float thickness_of(uint32_t R) { return 2.0f * sqrtf((float)R); }   // == (float)(2*sqrt((double)R)), R < 2^24

__device__ __forceinline__ void block_reduce_stats(double sum, double sum2, double mn, double mx, long long cnt,
                                                   StatsPartial *__restrict__ out) {
    __shared__ StatsPartial sm[32];
#pragma unroll
    for (int k = 16; k >= 1; k >>= 1) {
        sum += __shfl_xor_sync(0xFFFFFFFFu, sum, k);
        sum2 += __shfl_xor_sync(0xFFFFFFFFu, sum2, k);
        mn = fmin(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, k));
        mx = fmax(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, k));
        cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, k);
    }
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nthreads = blockDim.x * blockDim.y * blockDim.z;
    const int warp = tid >> 5, lane = tid & 31;
    if (lane == 0) { sm[warp].sum = sum; sm[warp].sum2 = sum2; sm[warp].mn = mn; sm[warp].mx = mx; sm[warp].count = cnt; }
    __syncthreads();
    if (tid == 0) {
        StatsPartial r = sm[0];
        for (int i = 1; i < (nthreads + 31) / 32; ++i) {
            r.sum += sm[i].sum; r.sum2 += sm[i].sum2; r.mn = fmin(r.mn, sm[i].mn); r.mx = fmax(r.mx, sm[i].mx);
            r.count += sm[i].count;
        }
        *out = r;
    }
}

// Output and statistics cover z in [z_lo, z_hi) of the local volume; `out` and nothing else is indexed
// relative to z_lo (a sharded caller passes its slab plus halo planes and keeps the core).
__global__ void __launch_bounds__(FTX *FTY) finish_kernel(const uint32_t *__restrict__ rsq,
                                                          const uint8_t *__restrict__ original, int w, int h, int d,
                                                          int z_lo, int z_hi, int fgval, int calibrate, float pw,
                                                          float *__restrict__ out, StatsPartial *__restrict__ partials) {
    __shared__ uint32_t R[HZ][HY][HX];
    __shared__ uint8_t inter[IZ][IY][IX];
    const int bx = blockIdx.x * FTX, by = blockIdx.y * FTY, bz = z_lo + blockIdx.z * FTZ;
    const int tid = threadIdx.x + FTX * threadIdx.y;
    const size_t wh = (size_t)w * h;
    for (int i = tid; i < HZ * HY * HX; i += FTX * FTY) {
        const int lx = i % HX, ly = (i / HX) % HY, lz = i / (HX * HY);
        const int gx = bx + lx - 2, gy = by + ly - 2, gz = bz + lz - 2;
        uint32_t v = OUTSIDE;
        if (gx >= 0 && gx < w && gy >= 0 && gy < h && gz >= 0 && gz < d) v = __ldg(rsq + gz * wh + (size_t)gy * w + gx);
        R[lz][ly][lx] = v;
    }
    __syncthreads();
    for (int i = tid; i < IZ * IY * IX; i += FTX * FTY) {
        const int lx = i % IX, ly = (i / IX) % IY, lz = i / (IX * IY);
        const uint32_t c = R[lz + 1][ly + 1][lx + 1];
        uint8_t f = 0;
        if (c != 0u && c != OUTSIDE) {
            bool all = true;
#pragma unroll
            for (int dz = 0; dz < 3; ++dz)
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) all = all && (R[lz + dz][ly + dy][lx + dx] != 0u);
            f = all ? 1 : 0;
        }
        inter[lz][ly][lx] = f;
    }
    __syncthreads();
    double sum = 0, sum2 = 0, mn = INFINITY, mx = -INFINITY;
    long long cnt = 0;
    const int x = bx + threadIdx.x, y = by + threadIdx.y;
    if (x < w && y < h) {
        const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll 1
        for (int tz = 0; tz < FTZ; ++tz) {
            const int z = bz + tz;
            if (z >= z_hi || z >= d) break;
            const uint32_t c = R[tz + 2][ty + 2][tx + 2];
            float v = 0.0f;
            if (c != 0u) {
                if (inter[tz + 1][ty + 1][tx + 1]) {
                    v = thickness_of(c);
                } else {
                    float s = 0.0f;
                    int n = 0;
#pragma unroll
                    for (int q = 0; q < 26; ++q) {
                        const int dx = NB26[q][0], dy = NB26[q][1], dz = NB26[q][2];
                        if (inter[tz + 1 + dz][ty + 1 + dy][tx + 1 + dx]) { s += thickness_of(R[tz + 2 + dz][ty + 2 + dy][tx + 2 + dx]); ++n; }
                    }
                    v = n > 0 ? s / (float)n : thickness_of(c);
                }
            }
            const size_t p = z * wh + (size_t)y * w + x;
            if (original && (int)__ldg(original + p) != fgval) v = 0.0f;
            if (calibrate) {
                if (v == 0.0f) v = __int_as_float(0x7FC00000);
                v = v * pw;
            }
            if (out) out[p - (size_t)z_lo * wh] = v;
            if (v == v) {   // not NaN
                const double dv = (double)v;
                sum += dv; sum2 += dv * dv; mn = fmin(mn, dv); mx = fmax(mx, dv); ++cnt;
            }
        }
    }
    const int bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    block_reduce_stats(sum, sum2, mn, mx, cnt, partials + bid);
}

__global__ void __launch_bounds__(1024) reduce_partials_kernel(const StatsPartial *__restrict__ partials, int n,
                                                               StatsPartial *__restrict__ result) {
    double sum = 0, sum2 = 0, mn = INFINITY, mx = -INFINITY;
    long long cnt = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const StatsPartial p = partials[i];
        sum += p.sum; sum2 += p.sum2; mn = fmin(mn, p.mn); mx = fmax(mx, p.mx); cnt += p.count;
    }
    block_reduce_stats(sum, sum2, mn, mx, cnt, result);
}

int finish_partials(int w, int h, int nz) {
    return ((w + FTX - 1) / FTX) * ((h + FTY - 1) / FTY) * ((nz + FTZ - 1) / FTZ);
}

int launch_finish(const uint32_t *rsq, const uint8_t *original, int w, int h, int d, int z_lo, int z_hi, int inverse,
                  int calibrate, float pixel_width, float *out, StatsPartial *partials, int npartials,
                  StatsPartial *result, cudaStream_t s) {
    dim3 grid((w + FTX - 1) / FTX, (h + FTY - 1) / FTY, (z_hi - z_lo + FTZ - 1) / FTZ);
    dim3 block(FTX, FTY, 1);
    finish_kernel<<<grid, block, 0, s>>>(rsq, original, w, h, d, z_lo, z_hi, inverse ? 0 : 255, calibrate, pixel_width, out,
                                         partials);
    reduce_partials_kernel<<<1, 1024, 0, s>>>(partials, npartials, result);
    return 2;
}

// ---- statistics of an existing float map ---------------------------------------------------------
__global__ void __launch_bounds__(256) stats_kernel(const float *__restrict__ map, size_t n,
                                                    StatsPartial *__restrict__ partials) {
    double sum = 0, sum2 = 0, mn = INFINITY, mx = -INFINITY;
    long long cnt = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = map[i];
        if (v == v) {
            const double dv = (double)v;
            sum += dv; sum2 += dv * dv; mn = fmin(mn, dv); mx = fmax(mx, dv); ++cnt;
        }
    }
    block_reduce_stats(sum, sum2, mn, mx, cnt, partials + blockIdx.x);
}

int launch_stats(const float *map, size_t n, StatsPartial *partials, int npartials, StatsPartial *result,
                 cudaStream_t s) {
    stats_kernel<<<npartials, 256, 0, s>>>(map, n, partials);
    reduce_partials_kernel<<<1, 1024, 0, s>>>(partials, npartials, result);
    return 2;
}

// ---- LocalThicknessWrapper.processImage input check: only 0 and 255 allowed ------------------------
__global__ void check_binary_kernel(const uint8_t *__restrict__ in, size_t n, uint32_t *__restrict__ bad) {
    uint32_t b = 0;
    const size_t n16 = ((uintptr_t)in % 16 == 0) ? n / 16 : 0;
    const uint4 *v = reinterpret_cast<const uint4 *>(in);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 q = __ldg(v + i);
        // a byte is fine iff it equals 0x00 or 0xFF  <=>  byte ^ (its top bit replicated) == 0
        const uint32_t wds[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t hi = (wds[k] >> 7) & 0x01010101u;
            b |= wds[k] ^ (hi * 0xFFu);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (size_t i = n16 * 16; i < n; ++i) b |= (in[i] != 0 && in[i] != 255);
    if (__any_sync(0xFFFFFFFFu, b != 0) && (threadIdx.x & 31) == 0) atomicOr(bad, 1u);
}

int launch_check_binary(const uint8_t *in, size_t n, uint32_t *bad, cudaStream_t s) {
    const unsigned grid = (unsigned)min((size_t)148 * 8, (n / 16 + 255) / 256);
    check_binary_kernel<<<grid ? grid : 1, 256, 0, s>>>(in, n, bad);
    return 1;
}

}  // namespace bjt
