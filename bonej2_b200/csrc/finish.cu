// finish.cu -- everything after painting, sm_100a:
//   (float)(2*sqrt(R))                      tail of Local_Thickness_Parallel            
//   Clean_Up_Local_Thickness                setFlag / averageInteriorNeighbors          (A.4)
//   MaskThicknessMapWithOriginal.trimOverhang                                            (A.5)
//   0 -> NaN, x (float)pixelWidth           LocalThicknessWrapper calibration tail       (A.6)
//   ij.process.StackStatistics              count / sum / sum^2 / min / max in double    (A.7)
// as driven by ThicknessWrapper.createMap / addMapResults (ThicknessWrapper.java:158-196).
//
// Two kernels.  flags: 0 = empty voxel, 1 = surface (non-zero with a zero 26-neighbour inside the
// image), 2 = interior.  final: interior voxels keep 2*sqrt(R); surface voxels take the float32
// mean of their interior neighbours, visited in the order faces, edges, corners (the order fixes
// the last ulp); then mask, NaN, calibration and the statistics partials in the same pass.
#include "common.cuh"
#include <math.h>

namespace bjt {

constexpr int FTX = 32, FTY = 8, FTZ = 4;

__global__ void __launch_bounds__(FTX *FTY *FTZ) flags_kernel(const uint32_t *__restrict__ rsq, int w, int h, int d,
                                                              uint8_t *__restrict__ flags) {
    __shared__ uint8_t nz[FTZ + 2][FTY + 2][FTX + 2];   // 1 = non-zero or outside the image
    const int bx = blockIdx.x * FTX, by = blockIdx.y * FTY, bz = blockIdx.z * FTZ;
    const int tid = threadIdx.x + FTX * (threadIdx.y + FTY * threadIdx.z);
    const size_t wh = (size_t)w * h;
    for (int i = tid; i < (FTZ + 2) * (FTY + 2) * (FTX + 2); i += FTX * FTY * FTZ) {
        const int lx = i % (FTX + 2), ly = (i / (FTX + 2)) % (FTY + 2), lz = i / ((FTX + 2) * (FTY + 2));
        const int gx = bx + lx - 1, gy = by + ly - 1, gz = bz + lz - 1;
        uint8_t v = 1;
        if (gx >= 0 && gx < w && gy >= 0 && gy < h && gz >= 0 && gz < d)
            v = __ldg(rsq + gz * wh + (size_t)gy * w + gx) != 0u;
        nz[lz][ly][lx] = v;
    }
    __syncthreads();
    const int x = bx + threadIdx.x, y = by + threadIdx.y, z = bz + threadIdx.z;
    if (x >= w || y >= h || z >= d) return;
    const int lx = threadIdx.x + 1, ly = threadIdx.y + 1, lz = threadIdx.z + 1;
    uint8_t f = 0;
    if (nz[lz][ly][lx]) {
        int all = 1;
#pragma unroll
        for (int dz = -1; dz <= 1; ++dz)
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) all &= nz[lz + dz][ly + dy][lx + dx];
        f = all ? 2 : 1;
    }
    flags[z * wh + (size_t)y * w + x] = f;
}

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

// Each block walks FZ consecutive z-tiles so the partial count stays small.
__global__ void __launch_bounds__(FTX *FTY *FTZ) final_kernel(const uint32_t *__restrict__ rsq,
                                                              const uint8_t *__restrict__ flags,
                                                              const uint8_t *__restrict__ original, int w, int h, int d,
                                                              int fgval, int calibrate, float pw, float *__restrict__ out,
                                                              StatsPartial *__restrict__ partials) {
    const int x = blockIdx.x * FTX + threadIdx.x, y = blockIdx.y * FTY + threadIdx.y, z = blockIdx.z * FTZ + threadIdx.z;
    const size_t wh = (size_t)w * h;
    double sum = 0, sum2 = 0, mn = INFINITY, mx = -INFINITY;
    long long cnt = 0;
    if (x < w && y < h && z < d) {
        const size_t p = z * wh + (size_t)y * w + x;
        const uint8_t f = flags[p];
        float v = 0.0f;
        if (f == 2) {
            v = thickness_of(__ldg(rsq + p));
        } else if (f == 1) {
            float s = 0.0f;
            int n = 0;
#pragma unroll
            for (int q = 0; q < 26; ++q) {
                const int xx = x + NB26[q][0], yy = y + NB26[q][1], zz = z + NB26[q][2];
                if (xx < 0 || xx >= w || yy < 0 || yy >= h || zz < 0 || zz >= d) continue;
                const size_t pq = zz * wh + (size_t)yy * w + xx;
                if (__ldg(flags + pq) == 2) { s += thickness_of(__ldg(rsq + pq)); ++n; }
            }
            v = n > 0 ? s / (float)n : thickness_of(__ldg(rsq + p));
        }
        if (original && (int)__ldg(original + p) != fgval) v = 0.0f;
        if (calibrate) {
            if (v == 0.0f) v = __int_as_float(0x7FC00000);
            v = v * pw;
        }
        if (out) out[p] = v;
        if (v == v) {   // not NaN
            const double dv = (double)v;
            sum = dv; sum2 = dv * dv; mn = dv; mx = dv; cnt = 1;
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

int finish_partials(int w, int h, int d) {
    return ((w + FTX - 1) / FTX) * ((h + FTY - 1) / FTY) * ((d + FTZ - 1) / FTZ);
}

int launch_finish(const uint32_t *rsq, const uint8_t *original, int w, int h, int d, int inverse, int calibrate,
                  float pixel_width, uint8_t *flags, float *out, StatsPartial *partials, int npartials,
                  StatsPartial *result, cudaStream_t s) {
    dim3 grid((w + FTX - 1) / FTX, (h + FTY - 1) / FTY, (d + FTZ - 1) / FTZ);
    dim3 block(FTX, FTY, FTZ);
    flags_kernel<<<grid, block, 0, s>>>(rsq, w, h, d, flags);
    final_kernel<<<grid, block, 0, s>>>(rsq, flags, original, w, h, d, inverse ? 0 : 255, calibrate, pixel_width, out,
                                        partials);
    reduce_partials_kernel<<<1, 1024, 0, s>>>(partials, npartials, result);
    return 3;
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
        // a byte is fine iff it equals 0x00 or 0xFF  <=>  byte ^ (byte >> 7 replicated) == 0
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
