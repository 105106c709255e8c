// ridge.cu -- distance-ridge extraction on the squared EDT, sm_100a.
//
// Replaces sc.fiji.localThickness.Distance_Ridge (run / createTemplate / scanCube).
// A foreground voxel p is dropped when an in-image 26-neighbour q (c = number of non-zero offset
// components) has sq(q) >= T_c[sq(p)], where T_c[r^2] is the smallest squared radius of a grid
// ball centred one step away (c components) that contains the grid ball of radius^2 r^2 at p.
// Works on the integer squared map directly (the Java code recovers the same integers from its
// float sqrt map via (int)(v*v+0.5f)).
#include "common.cuh"

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

// ---- distinct r^2 values ----------------------------------------------------------------------
__global__ void collect_values_kernel(const uint8_t *__restrict__ present, uint32_t vmax, uint32_t *__restrict__ values,
                                      uint32_t *__restrict__ count) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = v <= vmax && v > 0 && present[v];
    const unsigned ballot = __ballot_sync(0xFFFFFFFFu, on);
    if (!ballot) return;
    const int lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(count, (uint32_t)__popc(ballot));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (on) values[base + __popc(ballot & ((1u << lane) - 1u))] = v;
}

int launch_collect_values(const uint8_t *present, uint32_t vmax, uint32_t *values, uint32_t *count, cudaStream_t s) {
    const unsigned grid = (unsigned)(((size_t)vmax + 1 + 255) / 256);
    collect_values_kernel<<<grid, 256, 0, s>>>(present, vmax, values, count);
    return 1;
}

__global__ void mark_present_kernel(const uint32_t *__restrict__ sq, size_t n, uint8_t *__restrict__ present,
                                    uint32_t *__restrict__ maxval) {
    uint32_t tmax = 0, prev = 0xFFFFFFFFu;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t v = sq[i];
        if (present && v != prev) { present[v] = 1; prev = v; }
        tmax = max(tmax, v);
    }
    tmax = __reduce_max_sync(0xFFFFFFFFu, tmax);
    if ((threadIdx.x & 31) == 0 && tmax) atomicMax(maxval, tmax);
}

int launch_mark_present(const uint32_t *sq, size_t n, uint8_t *present, uint32_t *maxval, cudaStream_t s) {
    const unsigned grid = (unsigned)min((size_t)148 * 16, (n + 255) / 256);
    mark_present_kernel<<<grid ? grid : 1, 256, 0, s>>>(sq, n, present, maxval);
    return 1;
}

// ---- template (scanCube): one block per distinct value ------------------------------------------
// T_c[r^2] = max over k,j >= 0 with k^2+j^2 <= r^2 of (k+dz)^2 + (j+dy)^2 + (floor(sqrt(r^2-k^2-j^2))+dx)^2
template <bool COMPACT>
__global__ void __launch_bounds__(128) ridge_template_kernel(const uint32_t *__restrict__ values, uint32_t nvalues,
                                                             uint32_t *__restrict__ t1, uint32_t *__restrict__ t2,
                                                             uint32_t *__restrict__ t3) {
    const uint32_t vi = blockIdx.x;
    if (vi >= nvalues) return;
    const uint32_t rsq = values ? values[vi] : vi + 1u;      // no list: every r^2 in 1..nvalues
    const uint32_t r = isqrt_floor(rsq);      // k, j beyond floor(sqrt) never satisfy k^2+j^2 <= r^2
    const uint32_t side = r + 1;
    uint32_t m1 = 0, m2 = 0, m3 = 0;
    for (uint32_t idx = threadIdx.x; idx < side * side; idx += blockDim.x) {
        const uint32_t k = idx / side, j = idx % side;
        const uint32_t kj = k * k + j * j;
        if (kj <= rsq) {
            const uint32_t i1 = isqrt_floor(rsq - kj) + 1u;          // dx = 1 in all three templates
            const uint32_t i1sq = i1 * i1;
            m1 = max(m1, k * k + j * j + i1sq);                       // (1,0,0)
            m2 = max(m2, k * k + (j + 1u) * (j + 1u) + i1sq);         // (1,1,0)
            m3 = max(m3, (k + 1u) * (k + 1u) + (j + 1u) * (j + 1u) + i1sq);   // (1,1,1)
        }
    }
    __shared__ uint32_t sm[3];
    if (threadIdx.x < 3) sm[threadIdx.x] = 0;
    __syncthreads();
    m1 = __reduce_max_sync(0xFFFFFFFFu, m1);
    m2 = __reduce_max_sync(0xFFFFFFFFu, m2);
    m3 = __reduce_max_sync(0xFFFFFFFFu, m3);
    if ((threadIdx.x & 31) == 0) { atomicMax(&sm[0], m1); atomicMax(&sm[1], m2); atomicMax(&sm[2], m3); }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t o = COMPACT ? vi : rsq;
        t1[o] = sm[0]; t2[o] = sm[1]; t3[o] = sm[2];
    }
}

int launch_ridge_template(const uint32_t *values, uint32_t nvalues, uint32_t *t1, uint32_t *t2, uint32_t *t3,
                          cudaStream_t s) {
    if (!nvalues) return 0;
    ridge_template_kernel<false><<<nvalues, 128, 0, s>>>(values, nvalues, t1, t2, t3);
    return 1;
}
int launch_ridge_template_compact(const uint32_t *values, uint32_t nvalues, uint32_t *t1, uint32_t *t2,
                                  uint32_t *t3, cudaStream_t s) {
    if (!nvalues) return 0;
    ridge_template_kernel<true><<<nvalues, 128, 0, s>>>(values, nvalues, t1, t2, t3);
    return 1;
}

// ---- ridge test ------------------------------------------------------------------------------------
// A voxel p is dropped when a face / edge / corner neighbour q has sq(q) >= T1 / T2 / T3 [sq(p)], i.e. when
//   max over the 6 face neighbours >= T1  or  max over the 12 edge neighbours >= T2  or  max over the 8
//   corner neighbours >= T3.
// Block = 32 x 8 columns, each thread walks RTZ planes of its column.  The (34 x 10 x (RTZ+2)) halo tile of
// sq sits in shared memory (whole tile rows loaded by a warp; out-of-image cells hold 0, which never
// reaches a threshold since T_c > r^2 >= 1).  Per plane a thread reduces its 3 x 3 in-plane neighbourhood
// to (centre A, max of the 4 in-plane face neighbours B, max of the 4 in-plane diagonals C); the three
// maxima of plane z then follow from the (A, B, C) of planes z-1, z, z+1:
//   face = max(B[z], A[z-1], A[z+1])   edge = max(C[z], B[z-1], B[z+1])   corner = max(C[z-1], C[z+1]).
constexpr int RTX = 32, RTY = 8, RTZ = 8;
constexpr int RHX = RTX + 2, RHY = RTY + 2, RHZ = RTZ + 2;

// the test itself on a staged halo tile [RHZ][RHY][PITCH] (cell (0, 0, 0) = voxel (bx - XO, by - 1, bz - 1), out-of-image cells 0)
template <int PITCH, int XO>
__device__ __forceinline__ void ridge_tile_test(const uint32_t *__restrict__ tile, int bx, int by, int bz, int w, int h, int d,
                                                const uint32_t *__restrict__ t1, const uint32_t *__restrict__ t2,
                                                const uint32_t *__restrict__ t3, uint32_t *__restrict__ ridge,
                                                unsigned long long *__restrict__ count) {
    const int tid = threadIdx.x + RTX * threadIdx.y;
    const int lane = tid & 31;
    const size_t wh = (size_t)w * h;
    const int x = bx + threadIdx.x, y = by + threadIdx.y;
    const int lx = threadIdx.x + XO, ly = threadIdx.y + 1;
    unsigned nridge = 0;
    uint32_t A0, B0, C0, A1, B1, C1, A2, B2, C2;
    auto plane = [&](int lz, uint32_t &A, uint32_t &B, uint32_t &C) {
        const uint32_t *p = tile + (lz * RHY + ly) * PITCH + lx;
        A = p[0];
        B = max(max(p[-1], p[1]), max(p[-PITCH], p[PITCH]));
        C = max(max(p[-PITCH - 1], p[-PITCH + 1]), max(p[PITCH - 1], p[PITCH + 1]));
    };
    plane(0, A0, B0, C0);
    plane(1, A1, B1, C1);
#pragma unroll
    for (int tz = 0; tz < RTZ; ++tz) {
        plane(tz + 2, A2, B2, C2);
        const int z = bz + tz;
        if (x < w && y < h && z < d) {
            const uint32_t v = A1;
            bool is_ridge = false;
            if (v > 0) {
                const uint32_t face = max(B1, max(A0, A2)), edge = max(C1, max(B0, B2)), corner = max(C0, C2);
                is_ridge = !(face >= __ldg(t1 + v) || edge >= __ldg(t2 + v) || corner >= __ldg(t3 + v));
            }
            ridge[z * wh + (size_t)y * w + x] = is_ridge ? v : 0u;
            nridge += is_ridge ? 1u : 0u;
        }
        A0 = A1; B0 = B1; C0 = C1;
        A1 = A2; B1 = B2; C1 = C2;
    }
    // block count of ridge points: one atomic per block (a per-warp atomic on one address serialises in L2)
    __shared__ unsigned int block_count;
    if (tid == 0) block_count = 0;
    __syncthreads();
    nridge = __reduce_add_sync(0xFFFFFFFFu, nridge);
    if (lane == 0 && nridge) atomicAdd(&block_count, nridge);
    __syncthreads();
    if (tid == 0 && block_count) atomicAdd(count, (unsigned long long)block_count);
}

__global__ void __launch_bounds__(RTX *RTY, 8) ridge_kernel(const uint32_t *__restrict__ sq, int w, int h, int d,
                                                         const uint32_t *__restrict__ t1,
                                                         const uint32_t *__restrict__ t2,
                                                         const uint32_t *__restrict__ t3,
                                                         uint32_t *__restrict__ ridge,
                                                         unsigned long long *__restrict__ count) {
    __shared__ uint32_t tile[RHZ][RHY][RHX];
    const int bx = blockIdx.x * RTX, by = blockIdx.y * RTY, bz = blockIdx.z * RTZ;
    const int tid = threadIdx.x + RTX * threadIdx.y;
    const int lane = tid & 31, warp = tid >> 5;
    const size_t wh = (size_t)w * h;
    constexpr int NWARP = (RTX * RTY) / 32, NROW = RHZ * RHY;
    const int gx0 = bx + lane - 1, gx1 = bx + lane + 31;
    // ROWS_IN_FLIGHT tile rows are requested before the first is stored: with one row per trip the warp waited a
    // full memory latency per 128 bytes (40 % of the kernel's samples sat on the stores to shared memory)
    constexpr int ROWS_IN_FLIGHT = 4;
    for (int r0 = warp; r0 < NROW; r0 += ROWS_IN_FLIGHT * NWARP) {
        uint32_t va[ROWS_IN_FLIGHT], vb[ROWS_IN_FLIGHT];
#pragma unroll
        for (int u = 0; u < ROWS_IN_FLIGHT; ++u) {      // straight-line code: predicated loads, no branch between them
            const int r = r0 + u * NWARP;
            const int rr = r < NROW ? r : 0;
            const int lz = rr / RHY, ly = rr % RHY;
            const int gy = by + ly - 1, gz = bz + lz - 1;
            const bool row_in = r < NROW && gy >= 0 && gy < h && gz >= 0 && gz < d;
            const uint32_t *src = sq + (row_in ? gz * wh + (size_t)gy * w : 0);
            va[u] = (row_in && gx0 >= 0 && gx0 < w) ? __ldg(src + gx0) : 0u;
            vb[u] = (lane < 2 && row_in && gx1 < w) ? __ldg(src + gx1) : 0u;
        }
#pragma unroll
        for (int u = 0; u < ROWS_IN_FLIGHT; ++u) {
            const int r = r0 + u * NWARP;
            if (r < NROW) {
                const int lz = r / RHY, ly = r % RHY;
                tile[lz][ly][lane] = va[u];
                if (lane < 2) tile[lz][ly][32 + lane] = vb[u];
            }
        }
    }
    __syncthreads();
    ridge_tile_test<RHX, 1>(&tile[0][0][0], bx, by, bz, w, h, d, t1, t2, t3, ridge, count);
}

#if BJT_USE_TMA
// The same with the halo tile brought in by the TMA: the (34 x 10 x 10) neighbourhood of the block is inside one box
// of the [d][h][w] volume at coordinates (bx - 4, by - 1, bz - 1), 40 wide: a box has to START on a 16-byte boundary
// of its row (x = bx - 1 raises "illegal instruction": tools/study/tma_box_test.cu) and be a multiple of 16 bytes
// wide.  Cells outside the volume arrive as zeros, which is the reference's rule for neighbours outside the image.
// One cp.async.bulk.tensor instruction replaces the staging loop (row index arithmetic, bounds predicates, 2 loads
// and 2 shared-memory stores per row and warp); eight blocks per SM keep eight tiles in flight.
constexpr int RHXT = 40, RXOT = 4;
__global__ void __launch_bounds__(RTX *RTY, 8) ridge_tma_kernel(const __grid_constant__ CUtensorMap map, int w, int h, int d,
                                                             const uint32_t *__restrict__ t1, const uint32_t *__restrict__ t2,
                                                             const uint32_t *__restrict__ t3, uint32_t *__restrict__ ridge,
                                                             unsigned long long *__restrict__ count) {
    __shared__ __align__(128) uint32_t tile[RHZ * RHY * RHXT];
    __shared__ uint64_t bar;
    const int bx = blockIdx.x * RTX, by = blockIdx.y * RTY, bz = blockIdx.z * RTZ;
    const int tid = threadIdx.x + RTX * threadIdx.y;
    if (tid == 0) {
        tma::mbar_init(&bar, 1);
        tma::mbar_expect_tx(&bar, (uint32_t)sizeof(tile));
        tma::load_3d(tile, &map, bx - RXOT, by - 1, bz - 1, &bar);
    }
    __syncthreads();                                     // the barrier is initialised before anybody polls it
    if (tid < 32) tma::mbar_wait(&bar, 0);               // one warp polls, the others sleep in the block barrier
    __syncthreads();
    ridge_tile_test<RHXT, RXOT>(tile, bx, by, bz, w, h, d, t1, t2, t3, ridge, count);
}
#endif

int launch_ridge(const uint32_t *sq, int w, int h, int d, const uint32_t *t1, const uint32_t *t2, const uint32_t *t3,
                 uint32_t *ridge, unsigned long long *count, cudaStream_t s) {
    dim3 grid((w + RTX - 1) / RTX, (h + RTY - 1) / RTY, (d + RTZ - 1) / RTZ);
    dim3 block(RTX, RTY, 1);
#if BJT_USE_TMA
    CUtensorMap map;
    if (tma::make_map_3d_u32(&map, sq, w, h, d, RHXT, RHY, RHZ)) {
        ridge_tma_kernel<<<grid, block, 0, s>>>(map, w, h, d, t1, t2, t3, ridge, count);
        return 1;
    }
#endif
    ridge_kernel<<<grid, block, 0, s>>>(sq, w, h, d, t1, t2, t3, ridge, count);
    return 1;
}

}  // namespace bjt
