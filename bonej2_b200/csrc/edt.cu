// edt.cu -- exact squared Euclidean distance transform (Saito-Toriwaki decomposition), sm_100a.
//
// Replaces sc.fiji.localThickness.EDT_S1D (Step1/Step2/Step3 threads), reached
// from ThicknessWrapper.createMap() (ThicknessWrapper.java:190-196).  Result per voxel:
//   sq(p) = min( min_{q background, q inside the image} |p-q|^2 , 3(n+1)^2 ),  background -> 0.
//
//   x pass  one warp per row.  The row's background flags live in registers as bit masks
//           (lane l owns 32 consecutive voxels of every 1024-voxel segment), nearest flags left
//           and right come from clz/ffs plus a warp prefix-max / suffix-min over the lanes, and
//           the in-row distance (not squared) goes out as u16: 1 B read + 2 B written per voxel.
//   y, z    a block stages TX adjacent lines completely in shared memory and every warp computes, for its
//           columns at once, h(t) = min_d f(t+d) + d^2 by walking d outwards on both sides until d^2 has
//           reached the running minimum of every lane (edt_tile_kernel): no candidate read leaves the SM,
//           HBM sees one read and one write per voxel.  Shapes that do not fit a tile fall back to one
//           thread per line (edt_line4_kernel / edt_line_kernel): the minimiser of f(y') + (y-y')^2 is
//           monotone in y, so the scan for y starts at the previous minimiser and stops as soon as
//           (y'-y)^2 reaches the running minimum; candidate reads hit L1.
#include "common.cuh"
#include <algorithm>

// the z pass stages its tiles with TMA bulk-tensor copies when the toolchain is nvcc (device code only exists there)
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

// ---------------------------------------------------------------------------------------------
// x pass

__device__ __forceinline__ uint32_t bytes_lt_mask4(uint32_t word, uint32_t t4) {
    // 4 bytes -> 4 bits (bit k = byte k < threshold)
    uint32_t cmp = __vcmpltu4(word, t4);
    return ((cmp & 0x01010101u) * 0x01020408u) >> 24;
}

__device__ __forceinline__ uint32_t mask16_of(uint4 v, uint32_t t4) {
    return bytes_lt_mask4(v.x, t4) | (bytes_lt_mask4(v.y, t4) << 4) | (bytes_lt_mask4(v.z, t4) << 8) |
           (bytes_lt_mask4(v.w, t4) << 12);
}

template <int NSEG, bool ALIGNED>
__global__ void __launch_bounds__(256) edt_x_kernel(const uint8_t *__restrict__ in, int w, long long rows,
                                                     int thresh, int inverse, uint16_t *__restrict__ gx) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;   // whole warp leaves together
    const uint8_t *r = in + row * (long long)w;
    uint16_t *o = gx + row * (long long)w;
    const int BIG = 1 << 30;

    uint32_t m[NSEG];
    const uint32_t t4 = (uint32_t)(thresh < 0 ? 0 : (thresh > 255 ? 255 : thresh)) * 0x01010101u;
#pragma unroll
    for (int s = 0; s < NSEG; ++s) {
        const int x0 = (s * 32 + lane) * 32;
        uint32_t bits = 0;
        if (x0 < w) {
            if (ALIGNED && x0 + 32 <= w && thresh >= 0 && thresh <= 255) {
                const uint4 *p = reinterpret_cast<const uint4 *>(r + x0);
                uint4 a = __ldg(p), b = __ldg(p + 1);
                bits = mask16_of(a, t4) | (mask16_of(b, t4) << 16);
            } else {
                const int nb = min(32, w - x0);
                for (int b = 0; b < nb; ++b) bits |= ((int)r[x0 + b] < thresh ? 1u : 0u) << b;
            }
            if (inverse) {
                const int nb = min(32, w - x0);
                const uint32_t valid = nb >= 32 ? 0xFFFFFFFFu : ((1u << nb) - 1u);
                bits = (~bits) & valid;
            }
        }
        m[s] = bits;
    }

    // exclusive prefix-max of "last background index" and suffix-min of "first background index"
    int cl[NSEG], cr[NSEG];
    int segL = -1;
#pragma unroll
    for (int s = 0; s < NSEG; ++s) {
        const int x0 = (s * 32 + lane) * 32;
        int v = m[s] ? x0 + 31 - __clz(m[s]) : -1;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            int u = __shfl_up_sync(0xFFFFFFFFu, v, k);
            if (lane >= k) v = max(v, u);
        }
        int ex = __shfl_up_sync(0xFFFFFFFFu, v, 1);
        if (lane == 0) ex = -1;
        cl[s] = max(segL, ex);
        segL = max(segL, __shfl_sync(0xFFFFFFFFu, v, 31));
    }
    int segR = BIG;
#pragma unroll
    for (int s = NSEG - 1; s >= 0; --s) {
        const int x0 = (s * 32 + lane) * 32;
        int v = m[s] ? x0 + __ffs(m[s]) - 1 : BIG;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            int u = __shfl_down_sync(0xFFFFFFFFu, v, k);
            if (lane + k < 32) v = min(v, u);
        }
        int ex = __shfl_down_sync(0xFFFFFFFFu, v, 1);
        if (lane == 31) ex = BIG;
        cr[s] = min(segR, ex);
        segR = min(segR, __shfl_sync(0xFFFFFFFFu, v, 0));
    }

    // every voxel on its own from the word's bits: nearest flag at or below bit b (highest set bit of the low part, else
    // the carry-in from the left), nearest at or above (lowest set bit of the high part, else the carry-in from the
    // right).  Eight voxels at a time, so the 32 distances of a lane never sit in registers together (the kernel ran at
    // 100 registers and 23 % occupancy with the sequential scan over an array of 32).
#pragma unroll
    for (int s = 0; s < NSEG; ++s) {
        const int x0 = (s * 32 + lane) * 32;
        if (x0 >= w) continue;
        const uint32_t bits = m[s];
        const int left = cl[s] >= 0 ? cl[s] : -BIG, right = cr[s];         // -BIG / BIG: no flag on that side
        const bool vec = ALIGNED && x0 + 32 <= w;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t dd[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int b = 8 * k + j;
                const uint32_t low = bits & (0xFFFFFFFFu >> (31 - b)), high = bits >> b;
                const int hl = low ? x0 + 31 - __clz(low) : left;
                const int hr = high ? x0 + b + __ffs(high) - 1 : right;
                const int x = x0 + b;
                const int dv = min(x - hl, hr - x);                         // both sides missing: >= BIG - x0 - 31, far above any extent
                dd[j] = dv >= (1 << 20) ? (uint32_t)GX_INF : (uint32_t)dv;
            }
            if (vec) {
                uint4 v;
                v.x = dd[0] | (dd[1] << 16); v.y = dd[2] | (dd[3] << 16); v.z = dd[4] | (dd[5] << 16); v.w = dd[6] | (dd[7] << 16);
                reinterpret_cast<uint4 *>(o + x0)[k] = v;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (x0 + 8 * k + j < w) o[x0 + 8 * k + j] = (uint16_t)dd[j];
            }
        }
    }
}

template <int NSEG>
static void edt_x_dispatch(const uint8_t *in, int w, long long rows, int thresh, int inverse, uint16_t *gx,
                           bool aligned, cudaStream_t s) {
    const int warps = 8;
    const unsigned grid = (unsigned)((rows + warps - 1) / warps);
    if (aligned)
        edt_x_kernel<NSEG, true><<<grid, warps * 32, 0, s>>>(in, w, rows, thresh, inverse, gx);
    else
        edt_x_kernel<NSEG, false><<<grid, warps * 32, 0, s>>>(in, w, rows, thresh, inverse, gx);
}

int launch_edt_x(const uint8_t *in, int w, int h, int d, int inverse, int thresh, uint16_t *gx, cudaStream_t s) {
    const long long rows = (long long)h * d;
    const bool aligned = (w % 16 == 0) && ((uintptr_t)in % 16 == 0) && ((uintptr_t)gx % 16 == 0);
    const int nseg = (w + 1023) / 1024;
    if (nseg <= 1) edt_x_dispatch<1>(in, w, rows, thresh, inverse, gx, aligned, s);
    else if (nseg <= 2) edt_x_dispatch<2>(in, w, rows, thresh, inverse, gx, aligned, s);
    else if (nseg <= 4) edt_x_dispatch<4>(in, w, rows, thresh, inverse, gx, aligned, s);
    else edt_x_dispatch<8>(in, w, rows, thresh, inverse, gx, aligned, s);
    return 1;
}

// ---------------------------------------------------------------------------------------------
// y / z passes

__device__ __forceinline__ uint32_t f_of(uint16_t v, uint32_t n0) { return v == GX_INF ? n0 : (uint32_t)v * (uint32_t)v; }
__device__ __forceinline__ uint32_t f_of(uint32_t v, uint32_t) { return v; }

// line index t -> first element (t / w) * outer + t % w ; successive line positions are `stride` apart.
template <typename TIn, bool MARK>
__global__ void __launch_bounds__(256) edt_line_kernel(const TIn *__restrict__ in, uint32_t *__restrict__ out, int w,
                                                        long long nlines, long long outer, long long stride, int L,
                                                        uint32_t n0, uint8_t *__restrict__ present,
                                                        uint32_t *__restrict__ maxval) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = t < nlines;
    uint32_t tmax = 0;
    if (valid) {
        const long long base = (t / w) * outer + (t % w);
        const TIn *col = in + base;
        uint32_t *oc = out + base;
        const uint32_t NONE = 0x7FFFFFFFu;
        int a = 0;                                  // leftmost minimiser of the previous position
        uint32_t fa = f_of(__ldg(col), n0);
        uint32_t prev = 0xFFFFFFFFu;
        for (int y = 0; y < L; ++y) {
            const uint32_t fy = (y == a) ? fa : f_of(__ldg(col + (long long)y * stride), n0);
            uint32_t res;
            if (fy == 0u) {                         // background of this run: distance 0, new minimiser
                res = 0u; a = y; fa = 0u;
            } else {
                const int da = y - a;
                uint32_t m = fa >= n0 ? NONE : fa + (uint32_t)(da * da);
                int best = a;
                uint32_t fbest = fa;
                for (int yp = a + 1; yp < L; ++yp) {
                    const int dy = yp - y;
                    if (dy > 0 && (uint32_t)(dy * dy) >= m) break;
                    const uint32_t fv = f_of(__ldg(col + (long long)yp * stride), n0);
                    if (fv < n0) {
                        const uint32_t v = fv + (uint32_t)(dy * dy);
                        if (v < m) { m = v; best = yp; fbest = fv; }
                    }
                }
                if (m == NONE) {                    // no finite candidate on the rest of the line
                    for (int yy = y; yy < L; ++yy) oc[(long long)yy * stride] = n0;
                    if (MARK && present) present[n0] = 1;
                    tmax = n0;
                    break;
                }
                res = m < n0 ? m : n0;
                a = best; fa = fbest;
            }
            oc[(long long)y * stride] = res;
            if (MARK) {
                if (present && res != prev) { present[res] = 1; prev = res; }
                tmax = max(tmax, res);
            }
        }
    }
    if (MARK) {
        tmax = __reduce_max_sync(0xFFFFFFFFu, tmax);
        if ((threadIdx.x & 31) == 0 && tmax) atomicMax(maxval, tmax);
    }
}

// ---- 4 columns per thread -------------------------------------------------------------------------------
// Same recurrence, four adjacent columns (x .. x+3) per thread in lockstep over the union of their
// candidate windows: one 8- or 16-byte load brings the candidates of all four, stores are 16 bytes, and
// the four independent minimum chains give the scheduler something to overlap.  Candidates in front of a
// column's own previous minimiser are strictly worse (monotone minimiser), so visiting them is harmless.
struct Vec4 { uint32_t v[4]; };
__device__ __forceinline__ Vec4 load4(const uint16_t *p, uint32_t n0) {
    const uint2 r = __ldg(reinterpret_cast<const uint2 *>(p));
    Vec4 o;
    o.v[0] = f_of((uint16_t)(r.x & 0xFFFFu), n0); o.v[1] = f_of((uint16_t)(r.x >> 16), n0);
    o.v[2] = f_of((uint16_t)(r.y & 0xFFFFu), n0); o.v[3] = f_of((uint16_t)(r.y >> 16), n0);
    return o;
}
__device__ __forceinline__ Vec4 load4(const uint32_t *p, uint32_t) {
    const uint4 r = __ldg(reinterpret_cast<const uint4 *>(p));
    Vec4 o;
    o.v[0] = r.x; o.v[1] = r.y; o.v[2] = r.z; o.v[3] = r.w;
    return o;
}

constexpr int PF = 8;      // prefetch distance (line positions)

template <typename TIn, bool MARK>
__global__ void __launch_bounds__(128) edt_line4_kernel(const TIn *__restrict__ in, uint32_t *__restrict__ out, int w4,
                                                         long long ngroups, long long outer, long long stride, int L,
                                                         uint32_t n0, uint8_t *__restrict__ present,
                                                         uint32_t *__restrict__ maxval) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = t < ngroups;
    uint32_t tmax = 0;
    if (valid) {
        const long long base = (t / w4) * outer + (t % w4) * 4;
        const TIn *col = in + base;
        uint32_t *oc = out + base;
        const uint32_t NONE = 0x7FFFFFFFu;
        int a[4];
        uint32_t fa[4], prev[4];
        bool dead[4];                                // no finite value on the rest of the line
        Vec4 cur = load4(col, n0);
#pragma unroll
        for (int c = 0; c < 4; ++c) { a[c] = 0; fa[c] = cur.v[c]; prev[c] = 0xFFFFFFFFu; dead[c] = false; }
        for (int y = 0; y < L; ++y) {
            const Vec4 fy = cur;
            if (y + 1 < L) cur = load4(col + (long long)(y + 1) * stride, n0);     // the next position, in registers
            if (y + PF < L)                                                        // and a few further ones into L1
                asm volatile("prefetch.global.L1 [%0];" ::"l"(col + (long long)(y + PF) * stride));
            uint32_t m[4], fbest[4];
            int best[4];
            uint32_t mmax = 0;
            int start = L;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                best[c] = a[c]; fbest[c] = fa[c];
                if (fy.v[c] == 0u) { m[c] = 0u; best[c] = y; fbest[c] = 0u; }
                else if (dead[c]) { m[c] = 0u; }
                else {
                    const int da = y - a[c];
                    m[c] = fa[c] >= n0 ? NONE : fa[c] + (uint32_t)(da * da);
                    start = min(start, a[c] + 1);
                }
                mmax = max(mmax, m[c]);
            }
            if (mmax) {
                for (int yp = start; yp < L; ++yp) {
                    const int dy = yp - y;
                    const uint32_t dy2 = (uint32_t)(dy * dy);
                    if (dy > 0 && dy2 >= mmax) break;
                    const Vec4 fv = load4(col + (long long)yp * stride, n0);
                    uint32_t nm = 0;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t cand = fv.v[c] + dy2;
                        if (fv.v[c] < n0 && cand < m[c]) { m[c] = cand; best[c] = yp; fbest[c] = fv.v[c]; }
                        nm = max(nm, m[c]);
                    }
                    mmax = nm;
                }
            }
            uint4 res;
            uint32_t r[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (fy.v[c] != 0u && !dead[c] && m[c] == NONE) dead[c] = true;      // scanned to the end, nothing finite
                r[c] = (fy.v[c] == 0u) ? 0u : (dead[c] ? n0 : (m[c] < n0 ? m[c] : n0));
                if (!dead[c]) { a[c] = best[c]; fa[c] = fbest[c]; }
                if (MARK) {
                    if (present && r[c] != prev[c]) { present[r[c]] = 1; prev[c] = r[c]; }
                    tmax = max(tmax, r[c]);
                }
            }
            res.x = r[0]; res.y = r[1]; res.z = r[2]; res.w = r[3];
            *reinterpret_cast<uint4 *>(oc + (long long)y * stride) = res;
        }
    }
    if (MARK) {
        tmax = __reduce_max_sync(0xFFFFFFFFu, tmax);
        if ((threadIdx.x & 31) == 0 && tmax) atomicMax(maxval, tmax);
    }
}

// ---- whole lines in shared memory ---------------------------------------------------------------------
// A block stages TX adjacent lines completely in shared memory ([L][TX] squared values, loaded as
// contiguous TX-element row pieces), then every warp takes line positions t and computes, for its TX
// columns at once, h(t) = min_d f(t+d) + d^2 by walking d = 1, 2, ... outwards on both sides until d^2
// has reached the running minimum of every lane (one vote per step).  No candidate read leaves the SM,
// nothing depends on a load from global memory inside the loop, and the results go out as contiguous
// row pieces straight from registers.  HBM traffic is exactly one read and one write per voxel.
constexpr int LOADS_IN_FLIGHT = 8;
constexpr int EDT_PAD = 64;     // sentinel rows on either side of the staged lines: walks up to this distance need no clamping

// the walk over a staged tile: sm[t * TX + xx], t in [-EDT_PAD, L + EDT_PAD), pad rows = BIG
// (K line positions per lane in this 32-bit form -- a candidate row serving K positions -- measured slower: the loads fall
// by K but not the add-then-min per pair, and the ALU pipe becomes the limit: profiles/r02_stage_edt_walk_variants.log.)
template <int TX, bool WANT_MAX>
__device__ __forceinline__ uint32_t edt_tile_walk(const uint32_t *sm, uint32_t *__restrict__ out, long long base, long long stride, int L,
                                                  uint32_t n0) {
    constexpr int RPW = 32 / TX;                         // line positions per warp step
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int xx = lane % TX, rsub = lane / TX;
    uint32_t tmax = 0;
    for (int r0 = warp * RPW; r0 < L; r0 += nwarps * RPW) {
        const int t = r0 + rsub;
        const bool valid = t < L;
        const int tc = valid ? t : 0;
        const uint32_t *pc = sm + tc * TX + xx;
        uint32_t m = valid ? *pc : 0u;
        // four distances per vote.  Up to EDT_PAD the pad rows stand in for "beyond the line": no clamping
        int dlt = 1;
        const uint32_t *lo = pc - TX, *hi = pc + TX;
        for (; dlt + 3 <= EDT_PAD; dlt += 4, lo -= 4 * TX, hi += 4 * TX) {
            const uint32_t d2 = (uint32_t)(dlt * dlt);
            if (!__any_sync(0xFFFFFFFFu, d2 < m)) break;
            const uint32_t e = 2u * (uint32_t)dlt;
            const uint32_t d2b = d2 + e + 1u, d2c = d2 + 2u * e + 4u, d2d = d2 + 3u * e + 9u;
            const uint32_t m0 = min(min(lo[0], hi[0]) + d2, min(lo[-TX], hi[TX]) + d2b);
            const uint32_t m1 = min(min(lo[-2 * TX], hi[2 * TX]) + d2c, min(lo[-3 * TX], hi[3 * TX]) + d2d);
            m = min(m, min(m0, m1));
        }
        if (dlt + 3 > EDT_PAD) {                         // long walks: positions beyond the pad clamp onto its outermost rows
            for (;; dlt += 2) {
                const uint32_t d2 = (uint32_t)(dlt * dlt);
                if (!__any_sync(0xFFFFFFFFu, d2 < m)) break;
                const uint32_t d2b = d2 + 2u * (uint32_t)dlt + 1u;
                const uint32_t a = sm[max(tc - dlt, -1) * TX + xx], b = sm[min(tc + dlt, L) * TX + xx];
                const uint32_t a2 = sm[max(tc - dlt - 1, -1) * TX + xx], b2 = sm[min(tc + dlt + 1, L) * TX + xx];
                m = min(m, min(min(a, b) + d2, min(a2, b2) + d2b));
            }
        }
        if (valid) {
            const uint32_t res = m < n0 ? m : n0;
            out[base + (long long)t * stride + xx] = res;
            if (WANT_MAX) tmax = max(tmax, res);
        }
    }
    return tmax;                                         // this lane's largest result (0 unless WANT_MAX)
}

__device__ __forceinline__ void edt_publish_max(uint32_t tmax, uint32_t *__restrict__ maxval) {
    tmax = __reduce_max_sync(0xFFFFFFFFu, tmax);
    if ((threadIdx.x & 31) == 0 && tmax) atomicMax(maxval, tmax);
}

// staging of a tile of TX adjacent lines as u32: LOADS_IN_FLIGHT independent loads per thread before the first is used
// -- with one load per trip a thread waits a full memory latency for every 2-4 bytes (the profile of that form showed
// 37 % of the kernel's samples on the store to shared memory waiting for its load).  Pad rows: BIG (beyond the line).
template <typename TIn, int TX>
__device__ __forceinline__ void edt_stage_tile(const TIn *__restrict__ in, long long base, long long stride, int L, uint32_t n0,
                                               uint32_t *sm) {
    constexpr int RPW = 32 / TX;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int xx = lane % TX, rsub = lane / TX;
    const int rstep = nwarps * RPW;
    const long long pstep = (long long)rstep * stride;
    const TIn *p = in + base + (long long)(warp * RPW + rsub) * stride + xx;
    uint32_t *q = sm + (warp * RPW + rsub) * TX + xx;
    for (int t0 = warp * RPW + rsub; t0 < L; t0 += LOADS_IN_FLIGHT * rstep, p += LOADS_IN_FLIGHT * pstep, q += LOADS_IN_FLIGHT * rstep * TX) {
        TIn v[LOADS_IN_FLIGHT];
#pragma unroll
        for (int u = 0; u < LOADS_IN_FLIGHT; ++u) v[u] = t0 + u * rstep < L ? __ldg(p + u * pstep) : TIn(0);
#pragma unroll
        for (int u = 0; u < LOADS_IN_FLIGHT; ++u)
            if (t0 + u * rstep < L) q[u * rstep * TX] = f_of(v[u], n0);
    }
}

template <typename TIn, int TX, bool WANT_MAX>
__global__ void __launch_bounds__(1024, 2) edt_tile_kernel(const TIn *__restrict__ in, uint32_t *__restrict__ out, int groups,
                                                         long long outer, long long stride, int L, uint32_t n0,
                                                         uint32_t *__restrict__ maxval) {
    extern __shared__ uint32_t sm_raw[];                 // [EDT_PAD + L + EDT_PAD][TX]: the pad rows hold BIG (beyond the line)
    const long long base = (long long)(blockIdx.x / groups) * outer + (long long)(blockIdx.x % groups) * TX;
    const uint32_t BIG = 0x3FFFFFFFu;
    uint32_t *sm = sm_raw + EDT_PAD * TX;                // sm[t * TX + xx], t in [-EDT_PAD, L + EDT_PAD)
    for (int i = threadIdx.x; i < EDT_PAD * TX; i += blockDim.x) { sm[-EDT_PAD * TX + i] = BIG; sm[L * TX + i] = BIG; }
    edt_stage_tile<TIn, TX>(in, base, stride, L, n0, sm);
    __syncthreads();
    const uint32_t tmax = edt_tile_walk<TX, WANT_MAX>(sm, out, base, stride, L, n0);
    if (WANT_MAX) edt_publish_max(tmax, maxval);
}

// ---- two columns per lane, 16 bits each (DPX) -----------------------------------------------------------
// The walk above is bound by the ALU pipe (one add-then-min per candidate and position, half rate) and the
// shared-memory loads.  Distances on real data are far below 256 voxels, so the squared values that can win fit 16
// bits: this kernel keeps the tile as pairs of 16-bit values (two adjacent columns per 32-bit word, values clamped
// to EDT16_CLAMP, pad rows = the clamp), a lane owns one word-column and K consecutive line positions, and every
// VIADDMNMX.U16x2 (__viaddmin_u16x2) serves two (candidate, position) pairs; a row read from shared memory serves
// 2 K.  A clamped candidate can only win where the result itself is >= the clamp, and the walk only ends where
// d^2 has reached the running minimum, so a walk that ends within the pad rows is exact.  Where one does not (a
// voxel more than ~60 line positions from every feature, or a line without features) the block redoes its tile
// with the 32-bit walk, half of its columns at a time in the same shared memory.
// Pad rows and block shape, measured at 1024^3 (profiles/r02_stage_edt_pair_kernel_variants.log): 64 pad rows and
// three blocks of 512 threads per SM (73.7 KB each) -- three tiles in flight per SM hide the strided staging loads of
// the z pass better than two blocks of 1024 with 96 pad rows (y 1.70 / 2.22 -> 1.51 / 2.08 ms, z 2.27 -> 2.08 ms).
#ifndef BJT_EDT16_PAD
#define BJT_EDT16_PAD 64
#endif
#ifndef BJT_EDT16_THREADS
#define BJT_EDT16_THREADS 512
#define BJT_EDT16_BLOCKS 3
#endif
constexpr int EDT16_PAD = BJT_EDT16_PAD;
constexpr uint32_t EDT16_CLAMP = 55000u;                 // + (EDT16_PAD)^2 stays below 2^16
static_assert(EDT16_CLAMP + (uint32_t)(EDT16_PAD + 3) * (EDT16_PAD + 3) < 65536u, "16-bit sums must not wrap");
constexpr uint32_t EDT16_CLAMP2 = EDT16_CLAMP | (EDT16_CLAMP << 16);

#ifndef BJT_EDT16_K
#define BJT_EDT16_K 3                                    // odd: the two half-warps of a step then read different banks
#endif

__device__ __forceinline__ uint32_t edt16_pack(uint32_t a, uint32_t b) { return min(a, EDT16_CLAMP) | (min(b, EDT16_CLAMP) << 16); }
// y pass: two in-row distances (u16, GX_INF = none: 0xFFFF^2 still fits 32 bits and clamps)
__device__ __forceinline__ uint32_t edt16_load_pair(const uint16_t *p) {
    const uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(p));
    const uint32_t a = v & 0xFFFFu, b = v >> 16;
    return edt16_pack(a * a, b * b);
}
// z pass: two squared distances
__device__ __forceinline__ uint32_t edt16_load_pair(const uint32_t *p) {
    const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
    return edt16_pack(v.x, v.y);
}

// TW words (2 TW columns) per tile row: sm[t * TW + xw], t in [-EDT16_PAD, L + EDT16_PAD).  *sat: some walk of this warp did not end inside the pad rows.
template <bool WANT_MAX, int K, int TW>
__device__ __forceinline__ uint32_t edt16_walk(const uint32_t *sm, uint32_t *__restrict__ out, long long base, long long stride, int L,
                                               uint32_t n0, bool *sat) {
    static_assert(K >= 3 && (K & 1), "K: odd (bank parity of the half-warps) and at least 3 (D[4] is carried)");
    constexpr int GPW = 32 / TW;                         // groups of K positions per warp step
    constexpr int FAST = EDT16_PAD - (K - 1);
    constexpr uint32_t ONE2 = 0x00010001u;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int xw = lane & (TW - 1), gsub = lane / TW;
    uint32_t tmax = 0;
    for (int r0 = warp * GPW * K; r0 < L; r0 += nwarps * GPW * K) {
        const int t0 = r0 + gsub * K;
        const int tc = t0 < L ? t0 : 0;                  // a group beyond the line walks nothing and writes nothing
        const uint32_t *pc = sm + tc * TW + xw;
        uint32_t f[K], m[K];
#pragma unroll
        for (int k = 0; k < K; ++k) f[k] = pc[k * TW];   // rows at or beyond L are pad rows (the clamp)
#pragma unroll
        for (int k = 0; k < K; ++k) {                    // the positions of the group are candidates of one another
            m[k] = f[k];
#pragma unroll
            for (int q = 0; q < K; ++q)
                if (q != k) m[k] = __viaddmin_u16x2(f[q], (uint32_t)((k - q) * (k - q)) * ONE2, m[k]);
            if (t0 + k >= L) m[k] = 0u;
        }
        int j = 1;
        uint32_t D0 = ONE2, e = 2u * ONE2;               // j^2 and 2 j in both halves, carried from ring to ring
        const uint32_t *lo = pc - TW, *hi = pc + K * TW;
        for (;; j += 4, e += 8u * ONE2, lo -= 4 * TW, hi += 4 * TW) {
            uint32_t mx = m[0];
#pragma unroll
            for (int k = 1; k < K; ++k) mx = __vmaxu2(mx, m[k]);
            // some half of some position still above j^2?  (positions in the middle of the group have seen candidates
            // one or two further out; testing them against j^2 too costs nothing: the walk advances four rings a time)
            if (!__any_sync(0xFFFFFFFFu, __vmaxu2(mx, D0) != D0)) break;
            if (j + 3 > FAST) { *sat = true; break; }
            uint32_t D[K + 3];                           // D[i] = (j + i)^2 = D[i - 1] + 2 j + (2 i - 1)
            D[0] = D0;
#pragma unroll
            for (int i = 1; i < K + 3; ++i) D[i] = D[i - 1] + e + (uint32_t)(2 * i - 1) * ONE2;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t a = lo[-u * TW], b = hi[u * TW];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    m[k] = __viaddmin_u16x2(a, D[u + k], m[k]);
                    m[k] = __viaddmin_u16x2(b, D[u + K - 1 - k], m[k]);
                }
            }
            D0 = D[4];
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (t0 + k < L) {
                uint2 res;
                res.x = min(m[k] & 0xFFFFu, n0);
                res.y = min(m[k] >> 16, n0);
                *reinterpret_cast<uint2 *>(out + base + (long long)(t0 + k) * stride + 2 * xw) = res;
                if (WANT_MAX) tmax = max(tmax, max(res.x, res.y));
            }
        }
    }
    return tmax;
}

template <typename TIn, bool WANT_MAX, int K, int TW>
__global__ void __launch_bounds__(BJT_EDT16_THREADS, BJT_EDT16_BLOCKS) edt_tile16_kernel(const TIn *__restrict__ in, uint32_t *__restrict__ out, int groups,
                                                           long long outer, long long stride, int L, uint32_t n0,
                                                           uint32_t *__restrict__ maxval) {
    extern __shared__ uint32_t sm_raw[];                 // [EDT16_PAD + L + EDT16_PAD][TW] pairs; the redo: [EDT_PAD + L + EDT_PAD][TW] u32
    __shared__ int s_sat;
    constexpr int GPW = 32 / TW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const long long base = (long long)(blockIdx.x / groups) * outer + (long long)(blockIdx.x % groups) * (2 * TW);
    uint32_t *sm = sm_raw + EDT16_PAD * TW;
    if (threadIdx.x == 0) s_sat = 0;
    for (int i = threadIdx.x; i < EDT16_PAD * TW; i += blockDim.x) { sm[-EDT16_PAD * TW + i] = EDT16_CLAMP2; sm[L * TW + i] = EDT16_CLAMP2; }
    {
        const int xw = lane & (TW - 1), rsub = lane / TW;
        const int rstep = nwarps * GPW;
        const long long pstep = (long long)rstep * stride;
        const TIn *p = in + base + (long long)(warp * GPW + rsub) * stride + 2 * xw;
        uint32_t *q = sm + (warp * GPW + rsub) * TW + xw;
        for (int t0 = warp * GPW + rsub; t0 < L; t0 += LOADS_IN_FLIGHT * rstep, p += LOADS_IN_FLIGHT * pstep, q += LOADS_IN_FLIGHT * rstep * TW) {
            uint32_t v[LOADS_IN_FLIGHT];
#pragma unroll
            for (int u = 0; u < LOADS_IN_FLIGHT; ++u) v[u] = t0 + u * rstep < L ? edt16_load_pair(p + u * pstep) : 0u;
#pragma unroll
            for (int u = 0; u < LOADS_IN_FLIGHT; ++u)
                if (t0 + u * rstep < L) q[u * rstep * TW] = v[u];
        }
    }
    __syncthreads();
    bool sat = false;
    uint32_t tmax = edt16_walk<WANT_MAX, K, TW>(sm, out, base, stride, L, n0, &sat);
    if (sat) s_sat = 1;
    __syncthreads();
    if (s_sat) {                                         // rare: the whole tile again, exact in 32 bits, TW columns at a time
        tmax = 0;
        uint32_t *sm32 = sm_raw + EDT_PAD * TW;              // [EDT_PAD + L + EDT_PAD][TW] u32: as many bytes as the pairs of 2 TW columns
        for (int half = 0; half < 2; ++half) {
            for (int i = threadIdx.x; i < EDT_PAD * TW; i += blockDim.x) { sm32[-EDT_PAD * TW + i] = 0x3FFFFFFFu; sm32[L * TW + i] = 0x3FFFFFFFu; }
            edt_stage_tile<TIn, TW>(in, base + half * TW, stride, L, n0, sm32);
            __syncthreads();
            tmax = max(tmax, edt_tile_walk<TW, WANT_MAX>(sm32, out, base + half * TW, stride, L, n0));
            __syncthreads();
        }
    }
    if (WANT_MAX) edt_publish_max(tmax, maxval);
}

// false: the shape does not fit (lines too long for two tiles per SM, width not a multiple of the tile's)
template <typename TIn, bool WANT_MAX, int TW>
static bool launch_edt_tile16_tw(const TIn *in, uint32_t *out, int w, long long nouter, long long outer, long long stride, int L,
                                 uint32_t n0, uint32_t *maxval, cudaStream_t s) {
    const size_t smem = (size_t)(L + 2 * std::max(EDT16_PAD, EDT_PAD)) * TW * 4;        // pairs of 2 TW columns; the redo's u32 tile of TW columns
    if (w % (2 * TW) != 0 || smem > (size_t)(227 * 1024 / BJT_EDT16_BLOCKS - 1024) || (uintptr_t)in % 8 != 0 || (uintptr_t)out % 8 != 0) return false;
    const int groups = w / (2 * TW);
    cudaFuncSetAttribute(edt_tile16_kernel<TIn, WANT_MAX, BJT_EDT16_K, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    edt_tile16_kernel<TIn, WANT_MAX, BJT_EDT16_K, TW><<<(unsigned)(nouter * groups), BJT_EDT16_THREADS, smem, s>>>(in, out, groups, outer, stride, L, n0,
                                                                                                  maxval);
    return true;
}
// 32 columns per tile (64- / 128-byte row pieces), 16 where the lines are too long for two such tiles per SM
template <typename TIn, bool WANT_MAX>
static bool launch_edt_tile16(const TIn *in, uint32_t *out, int w, long long nouter, long long outer, long long stride, int L,
                              uint32_t n0, uint32_t *maxval, cudaStream_t s) {
    return launch_edt_tile16_tw<TIn, WANT_MAX, 16>(in, out, w, nouter, outer, stride, L, n0, maxval, s) ||
           launch_edt_tile16_tw<TIn, WANT_MAX, 8>(in, out, w, nouter, outer, stride, L, n0, maxval, s);
}

#if BJT_USE_TMA
// The z pass with its tile brought in by the TMA: the [L][16] tile of 16 adjacent z lines is exactly a (16, 1, L) box of
// the [d][h][w] volume, copied into shared memory in boxes of BZ planes by one thread's cp.async.bulk.tensor
// instructions -- no load, address or store instruction per voxel in the staging phase (it was ~10 of the kernel's
// ~75 instructions per voxel, and the kernel is bound by instruction issue).  The other threads write the pad rows
// meanwhile; one warp then polls the tile's mbarrier while the rest sleep in the block barrier (with every warp
// polling, the block that waits takes issue slots from the block that walks: measured 4.7 ms against 2.7 ms).
// Two blocks share an SM: one walks while the other's copies are in flight.
// Measured and rejected (1024^3, both runs): a persistent block per SM with a ring of three tile buffers, two tiles
// always in flight: 3.2 / 3.2 ms against 2.7 / 3.1 -- the copies, not the walk, then bound the kernel: 64-byte rows
// are 1024 requests per tile, and with 8-column tiles (32-byte rows, room for a ring in each of two blocks) the time
// doubles to 6.2 ms: the TMA's rate here is per row, not per byte.  L2 promotion off costs the same factor.
template <bool WANT_MAX, int TX>
__global__ void __launch_bounds__(1024, 2) edt_tile_z_tma_kernel(const __grid_constant__ CUtensorMap map, uint32_t *__restrict__ out,
                                                               int groups, int w, long long plane, int L, int BZ, uint32_t n0,
                                                               uint32_t *__restrict__ maxval) {
    extern __shared__ __align__(128) uint32_t sm_raw[];
    __shared__ uint64_t bar;
    const int y = blockIdx.x / groups, x0 = (blockIdx.x % groups) * TX;
    const uint32_t BIG = 0x3FFFFFFFu;
    uint32_t *sm = sm_raw + EDT_PAD * TX;
    if (threadIdx.x == 0) tma::mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        tma::mbar_expect_tx(&bar, (uint32_t)L * TX * 4u);
        for (int z0 = 0; z0 < L; z0 += BZ) tma::load_3d(sm + z0 * TX, &map, x0, y, z0, &bar);
    }
    for (int i = threadIdx.x; i < EDT_PAD * TX; i += blockDim.x) { sm[-EDT_PAD * TX + i] = BIG; sm[L * TX + i] = BIG; }
    if (threadIdx.x < 32) tma::mbar_wait(&bar, 0);
    __syncthreads();
    const uint32_t tmax = edt_tile_walk<TX, WANT_MAX>(sm, out, (long long)y * w + x0, plane, L, n0);
    if (WANT_MAX) edt_publish_max(tmax, maxval);
}
#endif

// returns false when the shape does not fit (caller falls back to the per-thread line kernels)
template <typename TIn, bool WANT_MAX>
static bool launch_edt_tile(const TIn *in, uint32_t *out, int w, long long nouter, long long outer, long long stride,
                            int L, uint32_t n0, uint32_t *maxval, cudaStream_t s) {
    const size_t budget = 100 * 1024;      // two blocks per SM: one loads its tile while the other computes
    int tx = 32;
    while (tx > 4 && (size_t)(L + 2 * EDT_PAD) * tx * 4 > budget) tx >>= 1;
    if ((size_t)(L + 2 * EDT_PAD) * tx * 4 > budget || w % tx != 0) return false;
    const size_t smem = (size_t)(L + 2 * EDT_PAD) * tx * 4;
    const int groups = w / tx;
    const unsigned grid = (unsigned)(nouter * groups);
    const int threads = 1024;
#define BJT_LAUNCH_TILE(TXV)                                                                                        \
    {                                                                                                               \
        cudaFuncSetAttribute(edt_tile_kernel<TIn, TXV, WANT_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        edt_tile_kernel<TIn, TXV, WANT_MAX><<<grid, threads, smem, s>>>(in, out, groups, outer, stride, L, n0, maxval);     \
    }
    if (tx == 32) BJT_LAUNCH_TILE(32)
    else if (tx == 16) BJT_LAUNCH_TILE(16)
    else if (tx == 8) BJT_LAUNCH_TILE(8)
    else BJT_LAUNCH_TILE(4)
#undef BJT_LAUNCH_TILE
    return true;
}

#ifndef BJT_EDT16
#define BJT_EDT16 3             // bit 0: y pass, bit 1: z pass through the 16-bit pair kernel
#endif

int launch_edt_y(const uint16_t *gx, int w, int h, int d, uint32_t n0, uint32_t *gy, cudaStream_t s) {
    if ((BJT_EDT16 & 1) && launch_edt_tile16<uint16_t, false>(gx, gy, w, d, (long long)w * h, (long long)w, h, n0, nullptr, s)) return 1;
    if (launch_edt_tile<uint16_t, false>(gx, gy, w, d, (long long)w * h, (long long)w, h, n0, nullptr, s)) return 1;
    if (w % 4 == 0 && (uintptr_t)gx % 8 == 0 && (uintptr_t)gy % 16 == 0) {
        const long long ng = (long long)(w / 4) * d;
        edt_line4_kernel<uint16_t, false><<<(unsigned)((ng + 127) / 128), 128, 0, s>>>(gx, gy, w / 4, ng, (long long)w * h,
                                                                                      (long long)w, h, n0, nullptr, nullptr);
        return 1;
    }
    const long long nlines = (long long)w * d;
    const unsigned grid = (unsigned)((nlines + 255) / 256);
    edt_line_kernel<uint16_t, false><<<grid, 256, 0, s>>>(gx, gy, w, nlines, (long long)w * h, (long long)w, h, n0,
                                                          nullptr, nullptr);
    return 1;
}

#if BJT_USE_TMA
// false: this shape or driver cannot take the TMA kernel
template <int TX>
static bool launch_edt_z_tma_tx(const uint32_t *gy, int w, int h, int d, uint32_t n0, uint32_t *sq, uint32_t *maxval, cudaStream_t s) {
    const size_t smem = (size_t)(d + 2 * EDT_PAD) * TX * 4;
    if (w % TX != 0 || d % 64 != 0 || smem > 100 * 1024 || !maxval) return false;
    const int bz = d % 256 == 0 ? 256 : (d % 128 == 0 ? 128 : 64);         // whole boxes only: nothing may land in the pad rows
    CUtensorMap map;
    if (!tma::make_map_3d_u32(&map, gy, w, h, d, TX, 1, bz)) return false;
    cudaFuncSetAttribute(edt_tile_z_tma_kernel<true, TX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    edt_tile_z_tma_kernel<true, TX><<<(unsigned)((long long)h * (w / TX)), 1024, smem, s>>>(map, sq, w / TX, w, (long long)w * h, d, bz, n0,
                                                                                          maxval);
    return true;
}
// 32 lines per tile (128-byte rows: half the row requests per byte) where the lines are short enough for two such
// tiles per SM, else 16
static bool launch_edt_z_tma(const uint32_t *gy, int w, int h, int d, uint32_t n0, uint32_t *sq, uint32_t *maxval, cudaStream_t s) {
    return launch_edt_z_tma_tx<32>(gy, w, h, d, n0, sq, maxval, s) || launch_edt_z_tma_tx<16>(gy, w, h, d, n0, sq, maxval, s);
}
#endif

int launch_edt_z(const uint32_t *gy, int w, int h, int d, uint32_t n0, uint32_t *sq, uint8_t *present,
                 uint32_t *maxval, cudaStream_t s) {
    if ((BJT_EDT16 & 2) && !present && maxval &&
        launch_edt_tile16<uint32_t, true>(gy, sq, w, h, (long long)w, (long long)w * h, d, n0, maxval, s)) return 1;
#if BJT_USE_TMA
    if (!present && launch_edt_z_tma(gy, w, h, d, n0, sq, maxval, s)) return 1;
#endif
    if (!present && launch_edt_tile<uint32_t, true>(gy, sq, w, h, (long long)w, (long long)w * h, d, n0, maxval, s)) return 1;
    if (w % 4 == 0 && (uintptr_t)gy % 16 == 0 && (uintptr_t)sq % 16 == 0) {
        const long long ng = (long long)(w / 4) * h;
        edt_line4_kernel<uint32_t, true><<<(unsigned)((ng + 127) / 128), 128, 0, s>>>(gy, sq, w / 4, ng, (long long)w,
                                                                                     (long long)w * h, d, n0, present, maxval);
        return 1;
    }
    const long long nlines = (long long)w * h;
    const unsigned grid = (unsigned)((nlines + 255) / 256);
    edt_line_kernel<uint32_t, true><<<grid, 256, 0, s>>>(gy, sq, w, nlines, (long long)w, (long long)w * h, d, n0,
                                                         present, maxval);
    return 1;
}

}  // namespace bjt
