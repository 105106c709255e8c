// finish.cu -- everything after painting, sm_100a:
//   (float)(2*sqrt(R))                      tail of Local_Thickness_Parallel            
//   Clean_Up_Local_Thickness                setFlag / averageInteriorNeighbors          (A.4)
//   MaskThicknessMapWithOriginal.trimOverhang                                            (A.5)
//   0 -> NaN, x (float)pixelWidth           LocalThicknessWrapper calibration tail       (A.6)
//   ij.process.StackStatistics              count / sum / sum^2 / min / max in double    (A.7)
// as driven by ThicknessWrapper.createMap / addMapResults (ThicknessWrapper.java:158-196).
//
// One fused kernel.  A block owns 32 x 8 x 16 output voxels and stages 2*sqrt(R) of the (36 x 12 x 20)
// tile with its 2-voxel halo in shared memory as float (out-of-image cells hold 0 and are flagged "not zero",
// as the reference's look() returns -1 there); "not zero" / "interior" are kept as one 64-bit word per tile
// row, built with ballots.  From the tile it derives the "interior" flag (non-zero, no zero 26-neighbour) on
// the 1-voxel halo.  Surface voxels (non-zero, not interior, not removed by the mask) are few and scattered, so
// they are first collected in a list and then averaged with all lanes busy: each takes the float32 mean of its
// interior neighbours visited in the order faces, edges, corners (the order fixes the last ulp) and overwrites
// its own cell, which no other voxel reads (only interior cells are read as neighbours).  Then every voxel:
// mask, NaN, calibration, and per-block partials of the statistics.  HBM sees r^2 once (plus halo re-reads that
// hit L2), the original once and the map once: 9 B/voxel of actual traffic.
#include "common.cuh"
#include <math.h>

// the halo tile comes in by one TMA bulk-tensor copy when the toolchain is nvcc (device code only exists there)
#ifndef BJT_USE_TMA
#ifdef __CUDACC__
#define BJT_USE_TMA 1
#else
#define BJT_USE_TMA 0
#endif
#endif
#if BJT_USE_TMA
#include "tma.cuh"
#endif

namespace bjt {

constexpr int FTX = 32, FTY = 8, FTZ = 16;
constexpr int HX = FTX + 4, HY = FTY + 4, HZ = FTZ + 4;       // r^2 tile with 2-voxel halo
constexpr int IY = FTY + 2, IZ = FTZ + 2;                      // interior flag rows with 1-voxel halo
constexpr uint32_t OUTSIDE = 0xFFFFFFFFu;

// neighbour visiting order of the cleanup average (A.4): 6 faces, 12 edges, 8 corners -- see BJT_NBR below

__device__ __forceinline__ float thickness_of(uint32_t R) { return 2.0f * sqrtf((float)R); }   // == (float)(2*sqrt((double)R)), R < 2^24
// The same through a table of the first THICK_TAB values (filled by thickness_table_kernel with thickness_of itself,
// so the bits are identical): the correctly rounded square root is a dozen instructions, the kernels below are bound by
// instruction issue, and the few thousand distinct r^2 of a volume keep their entries in L1.
constexpr uint32_t THICK_TAB = 65536;
__device__ __forceinline__ float thickness_lut(uint32_t R, const float *__restrict__ tab) {
    return (tab && R < THICK_TAB) ? __ldg(tab + R) : thickness_of(R);
}
__global__ void thickness_table_kernel(float *__restrict__ tab) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < THICK_TAB) tab[i] = thickness_of(i);
}

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

// Shared state of a block: the tile of 2*sqrt(r^2) values (PITCH cells per row; cell (0, 0, 0) is voxel (bx - XO, by - 2,
// bz - 2)) and, one 64-bit word per tile row (z, y) with bit j = tile column j: cell is not zero (outside the image
// counts as not zero) / cell is interior ("inside the image" follows from the coordinates).
template <int PITCH>
struct FinishTile {
    float S[HZ * HY * PITCH];
    unsigned long long nzrow[HZ][HY], itrow[HZ][HY];
    uint32_t keptrow[FTZ][FTY];                 // bit x: the mask keeps the voxel (or there is no mask)
    uint16_t surf[(FTZ / 2) * FTY * FTX];       // surface voxels of half the block: tz << 8 | ty << 5 | tx
    int nsurf[2];
};

// everything after the tile has been staged (S, nzrow, inrow filled; a block barrier follows the mask rows here)
template <int PITCH, int XO>
__device__ __forceinline__ void finish_tile_rest(FinishTile<PITCH> &T, const uint8_t *__restrict__ original, int bx, int by, int bz,
                                                 int w, int h, int d, int z_lo, int z_hi, int fgval, int calibrate, float pw,
                                                 float *__restrict__ out, StatsPartial *__restrict__ partials) {
    const int tid = threadIdx.x + FTX * threadIdx.y;
    const int lane = tid & 31;
    const size_t wh = (size_t)w * h;
    auto S = [&](int z, int y, int x) -> float & { return T.S[(z * HY + y) * PITCH + x]; };
    // which of the block's voxels the mask keeps
    {
        const int x = bx + threadIdx.x, y = by + threadIdx.y;
        uint8_t o[FTZ];
#pragma unroll
        for (int tz = 0; tz < FTZ; ++tz) {
            const int z = bz + tz;
            o[tz] = (uint8_t)fgval;
            if (original && x < w && y < h && z < z_hi && z < d) o[tz] = __ldg(original + z * wh + (size_t)y * w + x);
        }
#pragma unroll
        for (int tz = 0; tz < FTZ; ++tz) {
            const unsigned m = __ballot_sync(0xFFFFFFFFu, (int)o[tz] == fgval);
            if (lane == 0) T.keptrow[tz][threadIdx.y] = m;
        }
    }
    __syncthreads();
    // interior(x) = inside && not zero && all 27 cells around it not zero: AND of the 9 neighbouring rows,
    // each already ANDed with its two x-shifts
    if (tid < IZ * IY) {
        const int iz = 1 + tid / IY, iy = 1 + tid % IY;
        unsigned long long all = ~0ull;
#pragma unroll
        for (int dz = -1; dz <= 1; ++dz)
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                const unsigned long long m = T.nzrow[iz + dz][iy + dy];
                all &= m & (m >> 1) & (m << 1);
            }
        // cells of the row that lie inside the image: tile columns [XO - bx, w - bx + XO) of rows with y, z in range
        const int gy = by + iy - 2, gz = bz + iz - 2;
        const int lo = max(0, XO - bx), hi = min(64, w - bx + XO);
        unsigned long long inside = 0ull;
        if (gy >= 0 && gy < h && gz >= 0 && gz < d && hi > lo)
            inside = (hi >= 64 ? ~0ull : ((1ull << hi) - 1ull)) & ~((1ull << lo) - 1ull);
        T.itrow[iz][iy] = all & inside;
    }
    __syncthreads();
    // surface voxels the mask keeps (a voxel the mask removes needs no value: the mask is applied after the cleanup,
    // so skipping its average changes nothing; its painted value still counts as "not zero" for its neighbours above)
    const int x = bx + threadIdx.x, y = by + threadIdx.y;
    const int cx = threadIdx.x + XO, cy = threadIdx.y + 2;
    const bool col_in = x < w && y < h;
    for (int half = 0; half < 2; ++half) {
#pragma unroll 4
        for (int tz = half * (FTZ / 2); tz < (half + 1) * (FTZ / 2); ++tz) {
            const int z = bz + tz;
            const bool is_surf = col_in && z < z_hi && z < d && ((T.keptrow[tz][threadIdx.y] >> threadIdx.x) & 1u) &&
                                 S(tz + 2, cy, cx) != 0.0f && !((T.itrow[tz + 2][cy] >> cx) & 1ull);
            const unsigned m = __ballot_sync(0xFFFFFFFFu, is_surf);
            if (m) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&T.nsurf[half], __popc(m));
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                if (is_surf) T.surf[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)(tz << 8 | threadIdx.y << 5 | threadIdx.x);
            }
        }
        __syncthreads();
        // their values: float32 mean of the interior neighbours, faces, edges, corners in turn.  The 9 flag rows sit
        // in registers; adding 0.0f for the other neighbours is exact.  The cell is overwritten in place.
        const int ns = T.nsurf[half];
        for (int i = tid; i < ns; i += FTX * FTY) {
            const int code = T.surf[i];
            const int sx = (code & 31) + XO, sy = ((code >> 5) & 7) + 2, sz = (code >> 8) + 2;
            unsigned long long rows[3][3];
#pragma unroll
            for (int dz = 0; dz < 3; ++dz)
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) rows[dz][dy] = T.itrow[sz + dz - 1][sy + dy - 1] >> (sx - 1);
            float s = 0.0f;
            int n = 0;
#define BJT_NBR(dx, dy, dz)                                                         \
    {                                                                              \
        const bool in_ = (rows[(dz) + 1][(dy) + 1] >> ((dx) + 1)) & 1ull;          \
        s += in_ ? S(sz + (dz), sy + (dy), sx + (dx)) : 0.0f;                      \
        n += in_ ? 1 : 0;                                                          \
    }
            BJT_NBR(0, 0, -1) BJT_NBR(0, 0, 1) BJT_NBR(0, -1, 0) BJT_NBR(0, 1, 0) BJT_NBR(-1, 0, 0) BJT_NBR(1, 0, 0)
            BJT_NBR(0, 1, -1) BJT_NBR(0, 1, 1) BJT_NBR(1, -1, 0) BJT_NBR(1, 1, 0) BJT_NBR(-1, 0, 1) BJT_NBR(1, 0, 1)
            BJT_NBR(0, -1, -1) BJT_NBR(0, -1, 1) BJT_NBR(-1, -1, 0) BJT_NBR(-1, 1, 0) BJT_NBR(-1, 0, -1) BJT_NBR(1, 0, -1)
            BJT_NBR(1, 1, 1) BJT_NBR(1, -1, 1) BJT_NBR(-1, 1, 1) BJT_NBR(-1, -1, 1)
            BJT_NBR(1, 1, -1) BJT_NBR(1, -1, -1) BJT_NBR(-1, 1, -1) BJT_NBR(-1, -1, -1)
#undef BJT_NBR
            if (n > 0) S(sz, sy, sx) = s / (float)n;
        }
        __syncthreads();
    }
    // sum and sum of squares in double (StackStatistics); smallest and largest as float (the values are floats: the same
    // numbers), NaN entries add 0.0 and are skipped by fminf / fmaxf -- no branch and no double compare per voxel
    double sum = 0, sum2 = 0;
    float mnf = INFINITY, mxf = -INFINITY;
    int cnt = 0;
    if (col_in) {
#pragma unroll 4
        for (int tz = 0; tz < FTZ; ++tz) {
            const int z = bz + tz;
            if (z >= z_hi || z >= d) break;
            float v = S(tz + 2, cy, cx);
            if (!((T.keptrow[tz][threadIdx.y] >> threadIdx.x) & 1u)) v = 0.0f;
            if (calibrate) {
                if (v == 0.0f) v = __int_as_float(0x7FC00000);
                v = v * pw;
            }
            if (out) out[(size_t)(z - z_lo) * wh + (size_t)y * w + x] = v;
            const bool num = v == v;                                  // not NaN
            const double dv = num ? (double)v : 0.0;
            sum += dv; sum2 += dv * dv; mnf = fminf(mnf, v); mxf = fmaxf(mxf, v); cnt += num;
        }
    }
    const int bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    block_reduce_stats(sum, sum2, (double)mnf, (double)mxf, (long long)cnt, partials + bid);
}

// Output and statistics cover z in [z_lo, z_hi) of the local volume; `out` and nothing else is indexed
// relative to z_lo (a sharded caller passes its slab plus halo planes and keeps the core).
__global__ void __launch_bounds__(FTX *FTY, 4) finish_kernel(const uint32_t *__restrict__ rsq,
                                                          const uint8_t *__restrict__ original, int w, int h, int d,
                                                          int z_lo, int z_hi, int fgval, int calibrate, float pw,
                                                          float *__restrict__ out, StatsPartial *__restrict__ partials,
                                                          const float *__restrict__ ttab) {
    __shared__ FinishTile<HX> T;
    const int bx = blockIdx.x * FTX, by = blockIdx.y * FTY, bz = z_lo + blockIdx.z * FTZ;
    const int tid = threadIdx.x + FTX * threadIdx.y;
    const int lane = tid & 31, warp = tid >> 5;
    const size_t wh = (size_t)w * h;
    if (tid < 2) T.nsurf[tid] = 0;
    // a warp loads whole tile rows (36 cells: lanes 0..31, then lanes 0..3) and ballots the masks; row r = warp +
    // NWARP * j of the HZ * HY rows, (lz, ly) carried along instead of divided out
    constexpr int NWARP = (FTX * FTY) / 32, RPW = (HZ * HY) / NWARP;      // 30 rows per warp
    static_assert(RPW * NWARP == HZ * HY, "tile rows must divide evenly over the warps");
    static_assert(NWARP < HY, "the row step must stay below one plane of rows");
    constexpr int BATCH = 10;                  // rows (x 2 loads) a lane keeps in flight
    static_assert(RPW % BATCH == 0, "batches must divide the rows of a warp");
    const int gx0 = bx + lane - 2, gx1 = bx + lane + 30;
    const bool x0_in = gx0 >= 0 && gx0 < w, x1_in = lane < 4 && gx1 < w;
    int lz = 0, ly = warp;                      // row of the next load
#pragma unroll 1
    for (int b = 0; b < RPW / BATCH; ++b) {
        uint32_t v0[BATCH], v1[BATCH];
        int lz2 = lz, ly2 = ly;
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {
            const int gy = by + ly2 - 2, gz = bz + lz2 - 2;
            const bool row_in = gy >= 0 && gy < h && gz >= 0 && gz < d;
            const uint32_t *src = rsq + (row_in ? gz * wh + (size_t)gy * w : 0);
            v0[k] = (row_in && x0_in) ? __ldg(src + gx0) : OUTSIDE;
            v1[k] = (row_in && x1_in) ? __ldg(src + gx1) : OUTSIDE;
            ly2 += NWARP;
            if (ly2 >= HY) { ly2 -= HY; ++lz2; }
        }
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {
            // outside cells: value 0, flagged "not zero"
            T.S[(lz * HY + ly) * HX + lane] = thickness_lut(v0[k] == OUTSIDE ? 0u : v0[k], ttab);
            if (lane < 4) T.S[(lz * HY + ly) * HX + 32 + lane] = thickness_lut(v1[k] == OUTSIDE ? 0u : v1[k], ttab);
            const unsigned nz0 = __ballot_sync(0xFFFFFFFFu, v0[k] != 0u), nz1 = __ballot_sync(0xFFFFFFFFu, lane < 4 && v1[k] != 0u);
            if (lane == 0) T.nzrow[lz][ly] = (unsigned long long)nz0 | ((unsigned long long)nz1 << 32);
            ly += NWARP;
            if (ly >= HY) { ly -= HY; ++lz; }
        }
    }
    finish_tile_rest<HX, 2>(T, original, bx, by, bz, w, h, d, z_lo, z_hi, fgval, calibrate, pw, out, partials);
}

#if BJT_USE_TMA
// The same with the r^2 halo tile brought in by the TMA: one (40 x 12 x 20) box at (bx - 4, by - 2, bz - 2) -- a box
// starts on a 16-byte boundary of its row, so the 2-voxel halo in x sits inside a 4-voxel one --, cells outside the
// volume arriving as zeros.  A warp then turns the rows into 2*sqrt(r^2) in place (the float takes the place of the
// integer) and ballots the flag words, with "inside the image" from the coordinates instead of a marker value.
// No global address, bounds predicate or load instruction per cell: the staging loop was half of the kernel's
// instructions, and the kernel is bound by instruction issue.
// The tile in place: r^2 -> 2*sqrt(r^2), and the "not zero" word of every tile row (cells outside the volume -- zeros
// from the copy -- are flagged "not zero", as the reference's look() returns -1 there).  The 40 cells of a tile row are
// 32 + 8: the 8 beyond column 32 are taken four rows per warp step (with a row per step they kept 8 of 32 lanes busy
// through a second table lookup, store and ballot: 11 % of the kernel's instructions); their flags are the upper word
// of the row's 64-bit mask.
template <bool ALL_IN>
__device__ __forceinline__ void finish_convert_tile(FinishTile<40> &T, const float *__restrict__ ttab, int bx, int by, int bz, int w,
                                                    int h, int d, int tid) {
    constexpr int HXT = 40, FXO = 4;
    constexpr int NWARP = (FTX * FTY) / 32;
    const int lane = tid & 31, warp = tid >> 5;
    uint32_t *raw = reinterpret_cast<uint32_t *>(T.S);
    uint32_t *nzw = reinterpret_cast<uint32_t *>(&T.nzrow[0][0]);      // row r = lz * HY + ly: low word nzw[2 r], high word nzw[2 r + 1]
    for (int i = tid; i < HZ * HY * (HXT - 32); i += FTX * FTY) {      // whole warps: 8 cells x 4 rows per warp
        const int r = i >> 3, j = i & 7;
        bool in1 = true;
        if (!ALL_IN) {
            const int lz = r / HY, ly = r - lz * HY;
            const int gy = by + ly - 2, gz = bz + lz - 2;
            in1 = gy >= 0 && gy < h && gz >= 0 && gz < d && bx - FXO + 32 + j < w;
        }
        const uint32_t v1 = raw[r * HXT + 32 + j];
        T.S[r * HXT + 32 + j] = thickness_lut(v1, ttab);
        const unsigned b = __ballot_sync(0xFFFFFFFFu, v1 != 0u || !in1);
        if (j == 0) nzw[2 * r + 1] = (b >> (lane & 24)) & 0xFFu;
    }
    const int gx0 = bx - FXO + lane;
    const bool x0_in = gx0 >= 0 && gx0 < w;
    for (int r = warp, lz = 0, ly = warp; r < HZ * HY; r += NWARP) {   // NWARP <= HY: (lz, ly) advance without a division
        bool in0 = true;
        if (!ALL_IN) {
            const int gy = by + ly - 2, gz = bz + lz - 2;
            in0 = x0_in && gy >= 0 && gy < h && gz >= 0 && gz < d;
            ly += NWARP;
            if (ly >= HY) { ly -= HY; ++lz; }
        }
        const uint32_t v0 = raw[r * HXT + lane];
        T.S[r * HXT + lane] = thickness_lut(v0, ttab);
        const unsigned nz0 = __ballot_sync(0xFFFFFFFFu, v0 != 0u || !in0);
        if (lane == 0) nzw[2 * r] = nz0;
    }
}

constexpr int HXT = 40, FXO = 4;
__global__ void __launch_bounds__(FTX *FTY, 4) finish_tma_kernel(const __grid_constant__ CUtensorMap map,
                                                              const uint8_t *__restrict__ original, int w, int h, int d,
                                                              int z_lo, int z_hi, int fgval, int calibrate, float pw,
                                                              float *__restrict__ out, StatsPartial *__restrict__ partials,
                                                              const float *__restrict__ ttab) {
    __shared__ __align__(128) FinishTile<HXT> T;         // S first: the copy lands on a 128-byte boundary
    __shared__ uint64_t bar;
    const int bx = blockIdx.x * FTX, by = blockIdx.y * FTY, bz = z_lo + blockIdx.z * FTZ;
    const int tid = threadIdx.x + FTX * threadIdx.y;
    const int lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        tma::mbar_init(&bar, 1);
        tma::mbar_expect_tx(&bar, (uint32_t)sizeof(T.S));
        tma::load_3d(T.S, &map, bx - FXO, by - 2, bz - 2, &bar);
    }
    if (tid < 2) T.nsurf[tid] = 0;
    __syncthreads();                                     // the barrier is initialised before anybody polls it
    if (tid < 32) tma::mbar_wait(&bar, 0);               // one warp polls, the others sleep in the block barrier
    __syncthreads();
    // blocks whose whole halo tile lies inside the volume (nine in ten at 1024^3) need no coordinates for the flags
    const bool all_in = bx >= FXO && bx - FXO + HXT <= w && by >= 2 && by - 2 + HY <= h && bz >= 2 && bz - 2 + HZ <= d;
    if (all_in) finish_convert_tile<true>(T, ttab, bx, by, bz, w, h, d, tid);
    else finish_convert_tile<false>(T, ttab, bx, by, bz, w, h, d, tid);
    finish_tile_rest<HXT, FXO>(T, original, bx, by, bz, w, h, d, z_lo, z_hi, fgval, calibrate, pw, out, partials);
}
#endif

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

int launch_thickness_table(float *tab, cudaStream_t s) {
    thickness_table_kernel<<<THICK_TAB / 256, 256, 0, s>>>(tab);
    return 1;
}
size_t thickness_table_bytes() { return (size_t)THICK_TAB * sizeof(float); }

int launch_finish(const uint32_t *rsq, const uint8_t *original, int w, int h, int d, int z_lo, int z_hi, int inverse,
                  int calibrate, float pixel_width, float *out, StatsPartial *partials, int npartials,
                  StatsPartial *result, const float *thick_tab, cudaStream_t s) {
    dim3 grid((w + FTX - 1) / FTX, (h + FTY - 1) / FTY, (z_hi - z_lo + FTZ - 1) / FTZ);
    dim3 block(FTX, FTY, 1);
    bool done = false;
#if BJT_USE_TMA
    CUtensorMap map;
    if (tma::make_map_3d_u32(&map, rsq, w, h, d, HXT, HY, HZ)) {
        cudaFuncSetAttribute(finish_tma_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        finish_tma_kernel<<<grid, block, 0, s>>>(map, original, w, h, d, z_lo, z_hi, inverse ? 0 : 255, calibrate, pixel_width, out, partials, thick_tab);
        done = true;
    }
#endif
    if (!done)
        finish_kernel<<<grid, block, 0, s>>>(rsq, original, w, h, d, z_lo, z_hi, inverse ? 0 : 255, calibrate, pixel_width, out,
                                             partials, thick_tab);
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
    // the bytes in front of the first 16-byte boundary (a slab of a larger buffer need not start on one) and behind the
    // last whole vector are checked one by one, a thread each; the aligned middle as 16-byte vectors
    const size_t head = min(n, (size_t)((16 - ((uintptr_t)in & 15)) & 15));
    const size_t n16 = (n - head) / 16;
    const uint4 *v = reinterpret_cast<const uint4 *>(in + head);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid < head) b |= (in[tid] != 0 && in[tid] != 255);
    if (tid < n - head - n16 * 16) { const uint8_t t = in[head + n16 * 16 + tid]; b |= (t != 0 && t != 255); }
    for (size_t i = tid; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 q = __ldg(v + i);
        // a byte is fine iff it equals 0x00 or 0xFF  <=>  byte ^ (its top bit replicated) == 0
        const uint32_t wds[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t hi = (wds[k] >> 7) & 0x01010101u;
            b |= wds[k] ^ (hi * 0xFFu);
        }
    }
    if (__any_sync(0xFFFFFFFFu, b != 0) && (threadIdx.x & 31) == 0) atomicOr(bad, 1u);
}

int launch_check_binary(const uint8_t *in, size_t n, uint32_t *bad, cudaStream_t s) {
    const unsigned grid = (unsigned)min((size_t)148 * 8, (n / 16 + 255) / 256);
    check_binary_kernel<<<grid ? grid : 1, 256, 0, s>>>(in, n, bad);
    return 1;
}

}  // namespace bjt
