// paint_disc.cu -- largest-inscribed-sphere painting as "fronts along z, then discs in the plane" (path 4), sm_100a.
//
// Problem (sc.fiji.localThickness.Local_Thickness_Parallel): out(p) = max{ R_q : |p-q|^2 <= R_q } over
// the ridge points q.  The ball test is split as  dz^2 <= R - (dx^2 + dy^2):
//   1. fronts (zfront_kernel).  A ridge point (x, y, z_q, R) reaches plane z of its own column with slack
//      rho = R - (z - z_q)^2 >= 0: inside plane z it is the disc of squared radius rho around (x, y), tagged R.  At a
//      voxel only the Pareto front of (rho, R) matters (more slack and a larger tag both cover more), so the kernel
//      emits per voxel the non-dominated pairs.  It is a gather: a block stages a column window of the ridge in shared
//      memory, a warp sorts the candidates of one column by tag, and its lanes -- one plane each -- walk them in
//      descending tag order keeping the best slack so far: every lane reads the same candidate (a broadcast), there is
//      no sweep state and no order dependence between voxels.
//   2. discs (disc_kernel).  out(x, y, z) = max{ R : (x-xc)^2 + (y-yc)^2 <= rho } over the items of plane z.  A block
//      owns a 64 x 32 tile of one plane in shared memory.  A disc is a set of row intervals; an interval [lo, hi] is
//      recorded with TWO atomicMax operations in a sparse table (level k = floor(log2(len)): entries lo and
//      hi - 2^k + 1 of level k each stand for 2^k cells), so a disc costs its number of rows, not its area, and the
//      rows yc + dy and yc - dy share everything but their address.  One pass per level pushes the table down to
//      cells, and the tile goes out in whole rows.  Warps take buckets of items from a shared counter; accepted items
//      wait in per-warp queues by size class until a full warp of rows can be painted at once.
// Integer max everywhere: the result does not depend on the order of anything and is bitwise the brute-force answer
// (paint_scatter.cu, the oracle's lto_paint).  Items live in a pool that is filled and consumed per chunk of planes,
// so the workspace is a few bytes per voxel of a chunk, not of the volume.
//
// z-slab sharded runs call the same code on slab + halo planes: fronts need the ridge within the largest radius of a
// plane (the north star's max-radius halo exchange), discs are plane-local.
#include "common.cuh"

#ifndef BJT_ZF_STAGE_UNROLL
#define BJT_ZF_STAGE_UNROLL 4          // loads in flight per warp while the ridge window is staged
#endif

namespace bjt {

namespace disc_detail {

constexpr int ZF_ZB = 32;              // planes per front block = lanes of a warp
#ifndef BJT_ZF_THREADS
#define BJT_ZF_THREADS 128
#endif
constexpr int ZF_THREADS = BJT_ZF_THREADS;
constexpr int ZF_WARPS = ZF_THREADS / 32;
constexpr int ZF_STAGE_UNROLL = BJT_ZF_STAGE_UNROLL;
constexpr int ZF_KEY_Z_BITS = 10;      // key = R << 10 | window position
#ifndef BJT_ZF_ALIAS
#define BJT_ZF_ALIAS 1
#endif
#ifndef BJT_DISC_PREFETCH
#define BJT_DISC_PREFETCH 1            // 1: the next 32 items of a bucket are loaded before the current 32 are painted
#endif
#ifndef BJT_DISC_LPT
#define BJT_DISC_LPT 1
#endif

constexpr int DTW = 64, DTH = 32;      // disc tile: 64 columns x 32 rows
constexpr int DT_STRIDE = DTW + 3;     // row stride of a table level: rows of a disc land in different banks
constexpr int DT_LEVELS = 6;           // levels 0..5 are full tables; level 6 (the whole row) is one entry per row
constexpr int DISC_THREADS = 512;
constexpr int DISC_WARPS = DISC_THREADS / 32;
constexpr int DISC_BL = 1024;          // buckets listed per round
constexpr int DISC_NCLS = 4;           // size classes: up to 4 / 8 / 16 / more values of |dy|
constexpr int DISC_QCAP = 40;          // entries per class queue: fewer than 8 left over + 32 new ones

// item: rho [0,28) | R [28,56) | x within its 32-column bucket [56,61)
__device__ __forceinline__ unsigned long long make_item(uint32_t rho, uint32_t R, int xl) {
    return (unsigned long long)rho | ((unsigned long long)R << 28) | ((unsigned long long)xl << 56);
}
// bucket descriptor: first item [0,36) | items [36,52) | floor(sqrt(largest rho)) [52,64)
__device__ __forceinline__ unsigned long long make_desc(unsigned long long off, uint32_t cnt, uint32_t maxe) {
    return off | ((unsigned long long)cnt << 36) | ((unsigned long long)maxe << 52);
}

// descending bitonic sort of 32 (TWO = false) or 64 keys held (key0 = element lane, key1 = element lane + 32) by one
// warp; fully unrolled, the direction bits fold into lane masks
template <bool TWO>
__device__ __forceinline__ void warp_sort_desc(uint32_t &k0, uint32_t &k1, int lane) {
#pragma unroll
    for (int k = 2; k <= (TWO ? 64 : 32); k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j == 32) {                        // partner is the other register of the same lane (k == 64)
                const uint32_t a = max(k0, k1), b = min(k0, k1);
                k0 = a; k1 = b;
            } else {
                // element i keeps the larger of the pair when (i & j) == 0 and its block sorts descending ((i & k) == 0),
                // or when it is the upper one of an ascending block
                const uint32_t o0 = __shfl_xor_sync(0xFFFFFFFFu, k0, j);
                const bool up0 = (((lane & j) == 0) == ((lane & k) == 0));
                k0 = up0 ? max(k0, o0) : min(k0, o0);
                if (TWO) {
                    const uint32_t o1 = __shfl_xor_sync(0xFFFFFFFFu, k1, j);
                    const bool up1 = (((lane & j) == 0) == (((lane + 32) & k) == 0));
                    k1 = up1 ? max(k1, o1) : min(k1, o1);
                }
            }
        }
    }
}

// descending bitonic sort of P (power of two) keys in shared memory by one warp
__device__ void smem_sort_desc(uint32_t *s, int P, int lane) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < P; i += 32) {
                const int p = i ^ j;
                if (p > i) {
                    const uint32_t a = s[i], b = s[p];
                    const bool desc = (i & k) == 0;
                    if (desc ? a < b : a > b) { s[i] = b; s[p] = a; }
                }
            }
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// fronts: block = 32 columns (x) x one row (y) x ZF_ZB planes.  ridge holds planes [0, dext); the chunk is
// planes [zc0, zc1); descriptors are indexed ((z - zc0) * h + y) * nxt + xt.
// One walk over the sorted candidates of a column for the plane of this lane.  EMIT = false: the number of items;
// EMIT = true: the items go to out[0 ..).  *best_out = the largest slack (-1: nothing reaches the plane).
template <bool EMIT>
__device__ __forceinline__ int front_walk(const uint32_t *__restrict__ keys, int n, int zwin, int xl,
                                          unsigned long long *__restrict__ out, int *best_out) {
    int best = -1, cnt = 0;
    uint32_t lastR = 0u;
    unsigned long long *o = out - 1;                                // EMIT: the slot of the last item written
    const unsigned long long xhi = (unsigned long long)xl << 56;
#pragma unroll 4
    for (int c = 0; c < n; ++c) {
        const uint32_t key = keys[c];                               // the same address in every lane
        const uint32_t R = key >> ZF_KEY_Z_BITS;
        const int dz = zwin - (int)(key & ((1u << ZF_KEY_Z_BITS) - 1u));
        const int s = (int)R - dz * dz;
        if (s > best) {
            best = s;
            if (R != lastR) { ++cnt; ++o; lastR = R; }              // same tag, more slack: replaces the previous item
            if (EMIT) *o = (unsigned long long)(uint32_t)s | ((unsigned long long)R << 28) | xhi;      // make_item(s, R, xl)
        }
    }
    *best_out = best;
    return cnt;
}

__global__ void __launch_bounds__(ZF_THREADS) zfront_kernel(const uint32_t *__restrict__ ridge, int w, int h, int dext,
                                                             int zc0, int zc1, int emax, int nxt,
                                                             unsigned long long *__restrict__ desc,
                                                             unsigned long long *__restrict__ pool,
                                                             unsigned long long pool_cap, unsigned long long *cursor,
                                                             uint32_t *flags) {
    extern __shared__ uint32_t zf_smem[];
    const int ZW = ZF_ZB + 2 * emax;
    int P = 32;
    while (P < ZW) P <<= 1;
#if BJT_ZF_ALIAS
    // one array for both: a column's candidates are compacted in place (entry n + rank <= the position it was read
    // from, all reads of a group of 32 precede its writes), so the window and the sorted keys share a row of P + 1
    // words (odd: the transposed staging stores of a warp land in 32 banks).  Half the shared memory of separate
    // arrays: the kernel's occupancy is bounded by it (6 blocks of 4 warps per SM before).
    const int ZWP = P + 1, KP = P + 1;
    uint32_t *lines = zf_smem;                              // [32][P + 1] column x of the block: window, then sorted candidates
    uint32_t *skeys = lines;
    int *ncand = (int *)(lines + 32 * KP);                  // [32]
#else
    const int ZWP = ZW | 1, KP = P;
    uint32_t *lines = zf_smem;                              // [32][ZWP]   column x of the block, window position
    uint32_t *skeys = lines + 32 * ZWP;                     // [32][P]     sorted candidates of every column
    int *ncand = (int *)(skeys + 32 * P);                   // [32]
#endif
    uint32_t *cntz = (uint32_t *)(ncand + 32);              // [ZF_ZB] items of the plane's bucket
    uint32_t *mez = cntz + ZF_ZB;                           // [ZF_ZB] largest slack + 1
    uint32_t *exz = mez + ZF_ZB;                            // [ZF_ZB] exclusive prefix of cntz
    unsigned long long *sbase = (unsigned long long *)(exz + ZF_ZB);   // [1] (8-byte aligned: the counts above are even)
    uint16_t *pc = (uint16_t *)(sbase + 1);                 // [ZF_ZB][32] items per voxel, then their exclusive prefix

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned bid = blockIdx.x;
    const int xt = (int)(bid % (unsigned)nxt);
    const int y = (int)((bid / (unsigned)nxt) % (unsigned)h);
    const int zs = (int)(bid / ((unsigned)nxt * (unsigned)h));
    const int z0 = zc0 + zs * ZF_ZB;                        // first plane of the block
    const int nz = min(ZF_ZB, zc1 - z0);
    const int x = xt * 32 + lane;

    if (threadIdx.x < ZF_ZB) mez[threadIdx.x] = 0u;
    // stage the window: planes z0 - emax .. z0 + ZF_ZB - 1 + emax of 32 columns, transposed to [column][plane]
#pragma unroll ZF_STAGE_UNROLL
    for (int zi = warp; zi < ZW; zi += ZF_WARPS) {
        const int gz = z0 - emax + zi;
        uint32_t v = 0u;
        if (gz >= 0 && gz < dext && x < w) v = __ldg(ridge + ((size_t)gz * h + y) * w + x);
        lines[lane * ZWP + zi] = v;
    }
    __syncthreads();

    // per column: candidates sorted by tag (descending), then the front of every plane counted (lane = plane)
    const int zwin = emax + lane;
    for (int line = warp; line < 32; line += ZF_WARPS) {
        uint32_t *mys = skeys + line * KP;
        int n = 0;
        for (int base = 0; base < ZW; base += 32) {
            const int zi = base + lane;
            const uint32_t R = zi < ZW ? lines[line * ZWP + zi] : 0u;
            // reaches a plane of the block?  distance from zi to window positions [emax, emax + nz - 1]
            const int dseg = zi < emax ? emax - zi : (zi > emax + nz - 1 ? zi - (emax + nz - 1) : 0);
            const bool ok = R != 0u && (uint32_t)(dseg * dseg) <= R;
            const unsigned b = __ballot_sync(0xFFFFFFFFu, ok);
            if (ok) mys[n + __popc(b & ((1u << lane) - 1u))] = (R << ZF_KEY_Z_BITS) | (uint32_t)zi;
            n += __popc(b);
        }
        __syncwarp();
        if (n > 1) {
            if (n <= 64) {
                uint32_t k0 = lane < n ? mys[lane] : 0u, k1 = lane + 32 < n ? mys[lane + 32] : 0u;
                if (n > 32) warp_sort_desc<true>(k0, k1, lane);
                else warp_sort_desc<false>(k0, k1, lane);
                if (lane < n) mys[lane] = k0;
                if (lane + 32 < n) mys[lane + 32] = k1;
            } else {
                int Pn = 128;
                while (Pn < n) Pn <<= 1;
                for (int i = n + lane; i < Pn; i += 32) mys[i] = 0u;
                __syncwarp();
                smem_sort_desc(mys, Pn, lane);
            }
            __syncwarp();
        }
        if (lane == 0) ncand[line] = n;
        int best;
        const int cnt = front_walk<false>(mys, n, zwin, line, nullptr, &best);
        pc[lane * 32 + line] = (uint16_t)(lane < nz ? cnt : 0);
        if (lane < nz && best >= 0) atomicMax(mez + lane, (uint32_t)best + 1u);
    }
    __syncthreads();

    // per plane: exclusive prefix of the voxel counts over the columns, total of the bucket
    for (int zl = warp; zl < ZF_ZB; zl += ZF_WARPS) {
        const uint32_t mine = pc[zl * 32 + lane];
        uint32_t inc = mine;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc += t;
        }
        pc[zl * 32 + lane] = (uint16_t)min(inc - mine, 65535u);
        if (lane == 31) cntz[zl] = inc;
    }
    __syncthreads();

    // one allocation for the block's buckets
    if (threadIdx.x == 0) {
        unsigned long long tot = 0;
        bool wide = false;
        for (int zl = 0; zl < nz; ++zl) { exz[zl] = (uint32_t)tot; tot += cntz[zl]; wide |= cntz[zl] > 65535u; }
        unsigned long long base = 0;
        if (tot) base = atomicAdd(cursor, tot);
        if (wide || base + tot > pool_cap) { atomicOr(flags, 1u); base = ~0ull; }
        *sbase = base;
    }
    __syncthreads();
    const unsigned long long base = *sbase;
    if (base == ~0ull) {                                    // pool exhausted: empty buckets, the flag makes the caller retry
        for (int zl = threadIdx.x; zl < nz; zl += ZF_THREADS)
            desc[((size_t)(z0 - zc0 + zl) * h + y) * nxt + xt] = 0ull;
        return;
    }
    for (int zl = threadIdx.x; zl < nz; zl += ZF_THREADS) {
        const uint32_t me = mez[zl] ? isqrt_fast(mez[zl] - 1u) : 0u;
        desc[((size_t)(z0 - zc0 + zl) * h + y) * nxt + xt] = cntz[zl] ? make_desc(base + exz[zl], cntz[zl], me) : 0ull;
    }

    // the same walk, writing: the items of voxel (line, plane) start at the bucket's base + the voxel's prefix
    if (lane < nz) {
        for (int line = warp; line < 32; line += ZF_WARPS) {
            int best;
            front_walk<true>(skeys + line * KP, ncand[line], zwin, line, pool + base + exz[lane] + pc[lane * 32 + line], &best);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// discs: block = one 64 x 32 tile of plane blockIdx.y of the chunk (grid.x = tiles of a plane)

// A warp paints rows for `nent` queue entries of one size class at once: G = 4 << CLS lanes per entry, lane `sub`
// of an entry takes |dy| = dlo + sub (+ G, ... in the last class) and paints the rows yc + dy and yc - dy that lie
// inside the tile.  Entry: rho, R, xc | yc << 16 (signed, relative to the tile), dlo | n << 16.
template <int CLS>
__device__ __forceinline__ void drain_pass(smem_addr_t Tsh, smem_addr_t T6sh, const uint4 *q, int first, int nent, int lane, int W, int H) {
    constexpr int G = 4 << CLS;
    const int ent = lane >> (2 + CLS), sub = lane & (G - 1);
    if (ent >= nent) return;
    const uint4 e = q[first + ent];
    const uint32_t rho = e.x, R = e.y;
    const int xcr = (int)(e.z << 16) >> 16, yyr = (int)e.z >> 16;
    const int dlo = (int)(e.w & 0xFFFFu), n = (int)(e.w >> 16);
    for (int q1 = sub; q1 < n; q1 += G) {
        const int dy = dlo + q1;
        const int a = isqrt_small(rho - (uint32_t)(dy * dy));
        const int lo = max(0, xcr - a), hi = min(W - 1, xcr + a);
        const int len = hi - lo + 1;                                // <= 0: the rows miss the tile's columns
        const int r1 = yyr + dy, r2 = yyr - dy;
        const bool in1 = len > 0 && (unsigned)r1 < (unsigned)H, in2 = len > 0 && dy != 0 && (unsigned)r2 < (unsigned)H;
        const int k = 31 - __clz(len | 1);
        // no branches: level >= DT_LEVELS is the whole-row table (one entry per row, no second entry); a reduction
        // that is off for this lane goes to the lane's scratch word instead
        const bool whole = k >= DT_LEVELS;
        const int o2 = whole ? 0 : len - (1 << k);                  // second entry: hi + 1 - 2^k, relative to lo
        const smem_addr_t b = whole ? T6sh : Tsh + (smem_addr_t)((k * (DTH * DT_STRIDE) + lo) * 4);
        const int rs = whole ? 4 : DT_STRIDE * 4;
        const smem_addr_t a1 = b + (smem_addr_t)(r1 * rs), a2 = b + (smem_addr_t)(r2 * rs);
        // scratch: 64 words behind the whole-row table (lane + o2 <= 62).  Where there is no second entry (o2 == 0) the
        // second reduction repeats the first: max is idempotent, and a select and a compare are saved
        const smem_addr_t off = T6sh + (smem_addr_t)((DTH + lane) * 4);
        const smem_addr_t t1 = in1 ? a1 : off, t2 = in2 ? a2 : off;
        smem_red_max(t1, R);
        smem_red_max(t1 + (smem_addr_t)(o2 * 4), R);
        smem_red_max(t2, R);
        smem_red_max(t2 + (smem_addr_t)(o2 * 4), R);
        if (CLS < DISC_NCLS - 1) break;                             // n <= G in every class but the last
    }
}

// the lanes whose item falls in class C append it to the warp's queue of that class; full passes are painted at once
template <int C>
__device__ __forceinline__ void enqueue(smem_addr_t T, smem_addr_t T6, uint4 *myq, int &qn, int cls, const uint4 &rec, int lane,
                                        int W, int H) {
    const unsigned m = __ballot_sync(0xFFFFFFFFu, cls == C);
    if (!m) return;
    uint4 *q = myq + C * DISC_QCAP;
    if (cls == C) q[qn + __popc(m & ((1u << lane) - 1u))] = rec;
    qn += __popc(m);
    __syncwarp();
    constexpr int E = 8 >> C;                               // entries one pass takes
    while (qn >= E) {
        qn -= E;
        drain_pass<C>(T, T6, q, qn, E, lane, W, H);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(DISC_THREADS, 2) disc_kernel(const unsigned long long *__restrict__ desc,
                                                                const unsigned long long *__restrict__ pool, int w, int h,
                                                                int nxt, int emax, int ntx, uint32_t *__restrict__ out) {
    extern __shared__ uint32_t dk_smem[];
    uint32_t *T = dk_smem;                                  // [DT_LEVELS][DTH][DT_STRIDE]
    uint32_t *T6 = T + DT_LEVELS * DTH * DT_STRIDE;         // [DTH]
    unsigned long long *bl_desc = (unsigned long long *)(T6 + DTH + 64);        // [DISC_BL]  (T6: [DTH] + 64 scratch words; offset is a multiple of 8 bytes)
    uint32_t *bl_aux = (uint32_t *)(bl_desc + DISC_BL);     // [DISC_BL] row | column bucket << 16
    uint4 *queues = (uint4 *)(bl_aux + DISC_BL);            // [DISC_WARPS][DISC_NCLS][DISC_QCAP]  (16-byte aligned)
    int *ctl = (int *)(queues + DISC_WARPS * DISC_NCLS * DISC_QCAP);       // [0] listed buckets, [1] next bucket
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tx = (int)(blockIdx.x % (unsigned)ntx), ty = (int)(blockIdx.x / (unsigned)ntx);
    const int zl = blockIdx.y;                              // plane within the chunk
    const int x0 = tx * DTW, y0 = ty * DTH;
    const int x1 = min(x0 + DTW, w) - 1, y1 = min(y0 + DTH, h) - 1;
    const int W = x1 - x0 + 1, H = y1 - y0 + 1;

    for (int i = threadIdx.x; i < DT_LEVELS * DTH * DT_STRIDE + DTH; i += DISC_THREADS) T[i] = 0u;

    uint4 *myq = queues + warp * (DISC_NCLS * DISC_QCAP);
    const smem_addr_t Tsh = smem_addr(T), T6sh = smem_addr(T6);
    int qn0 = 0, qn1 = 0, qn2 = 0, qn3 = 0;                 // entries waiting per class (the same in every lane)

    // buckets of the extended tile: rows ya..yb, column buckets ba..bb, listed DISC_BL descriptors at a time
    const int ya = max(0, y0 - emax), yb = min(h - 1, y1 + emax);
    const int ba = max(0, x0 - emax) >> 5, bb = min(nxt - 1, (x1 + emax) >> 5);
    const int nb = bb - ba + 1;
    const int ndesc = (yb - ya + 1) * nb;
    const unsigned long long *dplane = desc + (size_t)zl * h * nxt;
    for (int d0 = 0; d0 < ndesc; d0 += DISC_BL) {
        if (threadIdx.x == 0) { ctl[0] = 0; ctl[1] = 0; ctl[2] = 0; }
        __syncthreads();
        for (int i = d0 + threadIdx.x; i < min(ndesc, d0 + DISC_BL); i += DISC_THREADS) {
            const int yy = ya + i / nb, bx = ba + i % nb;
            const unsigned long long dsc = __ldg(dplane + (size_t)yy * nxt + bx);
            if (dsc) {                                      // can the bucket's largest disc touch the tile?
                const int me = (int)(dsc >> 52);
                const int dy = yy < y0 ? y0 - yy : (yy > y1 ? yy - y1 : 0);
                const int bx0 = bx * 32, bx1 = bx0 + 31;
                const int dx = bx1 < x0 ? x0 - bx1 : (bx0 > x1 ? bx0 - x1 : 0);
                if (dx * dx + dy * dy < (me + 1) * (me + 1)) {
#if BJT_DISC_LPT
                    // buckets with more than a warp of items are listed from the front, the others from the back: the
                    // warps take the long ones first and the round ends on short ones
                    const bool heavy = ((dsc >> 36) & 0xFFFFu) > 32u;
                    const int pos = heavy ? atomicAdd(ctl, 1) : DISC_BL - 1 - atomicAdd(ctl + 2, 1);
#else
                    const int pos = atomicAdd(ctl, 1);
#endif
                    bl_desc[pos] = dsc;
                    bl_aux[pos] = (uint32_t)yy | ((uint32_t)bx << 16);
                }
            }
        }
        __syncthreads();
#if BJT_DISC_LPT
        const int nheavy = ctl[0], nlist = ctl[0] + ctl[2];
#define BJT_DISC_SLOT(i) ((i) < nheavy ? (i) : DISC_BL - 1 - ((i) - nheavy))
#else
        const int nlist = ctl[0];
#define BJT_DISC_SLOT(i) (i)
#endif
        // Measured (1024^3 background run, profiles/r02_stage_painter_variants.log): taking the long buckets first saves
        // 1.9 .. 3.6 ms of 84.8, loading the next 32 items of a bucket before the current 32 are painted 1.9 more.
        // Claiming the NEXT BUCKET ahead and loading its first items early (10 % of the kernel's stall samples sat on
        // the first use of an item) costs 3.3 ms instead: a warp that sits on a claimed bucket while it paints
        // lengthens the tail of the round.  Dropping discs that lie inside a disc a few lanes away (same row, larger
        // tag): +8.6 ms for a window of 2 lanes either side -- every bucket is examined by ~9 tiles, and such a window
        // finds only 5 % of the disc rows (tools/study/lane_prune_study.py; 29 % with the whole bucket as window).
        while (true) {
            int bi = 0;
            if (lane == 0) bi = atomicAdd(ctl + 1, 1);
            bi = __shfl_sync(0xFFFFFFFFu, bi, 0);
            if (bi >= nlist) break;
            const unsigned long long bd = bl_desc[BJT_DISC_SLOT(bi)];
            const uint32_t aux = bl_aux[BJT_DISC_SLOT(bi)];
            const int byy = (int)(aux & 0xFFFFu), bxx = (int)(aux >> 16) * 32;
            const unsigned long long off = bd & ((1ull << 36) - 1ull);
            const int cnt = (int)((bd >> 36) & 0xFFFFu);
            const int dyl = byy < y0 ? y0 - byy : (byy > y1 ? byy - y1 : 0);
            const int yyr = byy - y0;
#if BJT_DISC_PREFETCH
            unsigned long long it = lane < cnt ? __ldg(pool + off + lane) : 0ull;
#endif
            for (int b0 = 0; b0 < cnt; b0 += 32) {
#if BJT_DISC_PREFETCH
                unsigned long long itn = 0ull;                  // the next 32 items of the bucket, before these are painted
                if (b0 + 32 + lane < cnt) itn = __ldg(pool + off + b0 + 32 + lane);
#else
                const unsigned long long it = b0 + lane < cnt ? __ldg(pool + off + b0 + lane) : 0ull;
#endif
                // lane = item: does its disc touch the tile, and over how many values of |dy|
                int cls = -1;
                uint4 rec = make_uint4(0u, 0u, 0u, 0u);
                if (b0 + lane < cnt) {
                    const uint32_t rho = (uint32_t)(it & 0xFFFFFFFu), R = (uint32_t)((it >> 28) & 0xFFFFFFFu);
                    const int xcr = bxx + (int)(it >> 56) - x0;
                    const int dxl = xcr < 0 ? -xcr : (xcr > W - 1 ? xcr - (W - 1) : 0);
                    const int rem = (int)rho - dxl * dxl;
                    if (rem >= dyl * dyl) {
                        // |dy| from dyl up to where the disc ends or both rows yc + dy, yc - dy have left the tile
                        const int n = min(isqrt_small((uint32_t)rem), max(yyr, H - 1 - yyr)) - dyl + 1;
                        cls = n <= 4 ? 0 : (n <= 8 ? 1 : (n <= 16 ? 2 : 3));
                        rec = make_uint4(rho, R, ((uint32_t)xcr & 0xFFFFu) | ((uint32_t)yyr << 16), (uint32_t)dyl | ((uint32_t)n << 16));
                    }
                }
                enqueue<0>(Tsh, T6sh, myq, qn0, cls, rec, lane, W, H);
                enqueue<1>(Tsh, T6sh, myq, qn1, cls, rec, lane, W, H);
                enqueue<2>(Tsh, T6sh, myq, qn2, cls, rec, lane, W, H);
                enqueue<3>(Tsh, T6sh, myq, qn3, cls, rec, lane, W, H);
#if BJT_DISC_PREFETCH
                it = itn;
#endif
            }
        }
        __syncthreads();                                    // the list is rewritten by the next round
    }
    // what is left in the queues
    if (qn0) drain_pass<0>(Tsh, T6sh, myq, 0, qn0, lane, W, H);
    if (qn1) drain_pass<1>(Tsh, T6sh, myq + DISC_QCAP, 0, qn1, lane, W, H);
    if (qn2) drain_pass<2>(Tsh, T6sh, myq + 2 * DISC_QCAP, 0, qn2, lane, W, H);
    if (qn3) drain_pass<3>(Tsh, T6sh, myq + 3 * DISC_QCAP, 0, qn3, lane, W, H);
    __syncthreads();

    // push the table down: level 6 (whole row) into the two halves of level 5, then level k into level k - 1
    for (int i = threadIdx.x; i < DTH * 2; i += DISC_THREADS) {
        const int row = i >> 1, j = (i & 1) * 32;
        uint32_t *p = T + ((DT_LEVELS - 1) * DTH + row) * DT_STRIDE + j;
        *p = max(*p, T6[row]);
    }
    __syncthreads();
    for (int k = DT_LEVELS - 1; k >= 1; --k) {
        const int half = 1 << (k - 1);
        const uint32_t *Lk = T + (k * DTH) * DT_STRIDE;
        uint32_t *Ll = T + ((k - 1) * DTH) * DT_STRIDE;
        for (int i = threadIdx.x; i < DTH * DTW; i += DISC_THREADS) {
            const int row = i >> 6, j = i & 63;
            uint32_t v = Lk[row * DT_STRIDE + j];
            if (j >= half) v = max(v, Lk[row * DT_STRIDE + j - half]);
            if (v) { uint32_t *p = Ll + row * DT_STRIDE + j; *p = max(*p, v); }
        }
        __syncthreads();
    }
    uint32_t *oplane = out + (size_t)zl * w * h;
    for (int i = threadIdx.x; i < DTH * DTW; i += DISC_THREADS) {
        const int row = i >> 6, j = i & 63;
        if (row < H && j < W) oplane[(size_t)(y0 + row) * w + x0 + j] = T[row * DT_STRIDE + j];
    }
}

// every argument isqrt_small is specified for, against the exact integer root
__global__ void isqrt_small_check_kernel(uint32_t *mismatches) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < (1u << 18) && (uint32_t)isqrt_small(v) != isqrt_floor(v)) atomicAdd(mismatches, 1u);
}

size_t zfront_smem_bytes(int emax) {
    const int ZW = ZF_ZB + 2 * emax;
    int P = 32;
    while (P < ZW) P <<= 1;
#if BJT_ZF_ALIAS
    const size_t rows = (size_t)32 * (P + 1);
#else
    const size_t rows = (size_t)32 * (ZW | 1) + (size_t)32 * P;
#endif
    return (rows + 32 + 3 * ZF_ZB) * 4 + 8 + (size_t)ZF_ZB * 32 * 2 + 16;
}
constexpr size_t DISC_SMEM = (size_t)(DT_LEVELS * DTH * DT_STRIDE + DTH + 64) * 4 + (size_t)DISC_BL * 12 +
                             (size_t)DISC_WARPS * DISC_NCLS * DISC_QCAP * 16 + 16;

uint32_t host_isqrt(uint64_t v) {
    uint32_t r = (uint32_t)sqrt((double)v);
    while ((uint64_t)r * r > v) --r;
    while ((uint64_t)(r + 1) * (r + 1) <= v) ++r;
    return r;
}

}  // namespace disc_detail
using namespace disc_detail;

// largest tag the disc painter takes: the front kernel's window must fit in shared memory, its keys in 32 bits
// (tag << 10 | window position), and the disc kernel's square roots are exact below 2^18
bool paint_disc_supports(uint32_t rsq_max) {
    if (rsq_max >= (1u << 18)) return false;                // isqrt_small; also keeps keys and queue entries in range
    const int emax = (int)host_isqrt(rsq_max);
    if (ZF_ZB + 2 * emax > (1 << ZF_KEY_Z_BITS)) return false;
    return zfront_smem_bytes(emax) <= 200 * 1024;
}

int launch_isqrt_small_check(uint32_t *mismatches, cudaStream_t s) {
    isqrt_small_check_kernel<<<(1u << 18) / 256, 256, 0, s>>>(mismatches);
    return 1;
}

// planes per chunk and the workspace for it: descriptors (8 B per 32 voxels of a chunk) + items_per_voxel * 8 B per
// voxel of a chunk + cursor.  budget_bytes bounds the workspace (0 = the default of 6 GB).
void paint_disc_plan(int w, int h, int nplanes, int items_per_voxel, size_t budget_bytes, int *chunk_planes,
                     size_t *workspace_bytes) {
    const size_t wh = (size_t)w * h;
    const int nxt = (w + 31) / 32;
    int zc = nplanes;
    const size_t target = budget_bytes ? budget_bytes : ((size_t)6 << 30);
    auto need = [&](int planes) {
        return (size_t)planes * h * nxt * 8 + (size_t)planes * wh * (size_t)items_per_voxel * 8 + 256;
    };
    while (zc > ZF_ZB && need(zc) > target) zc = ((zc / 2 + ZF_ZB - 1) / ZF_ZB) * ZF_ZB;
    *chunk_planes = zc;
    *workspace_bytes = need(zc);
}

// ridge: planes [0, dext) of w x h; paints planes [z_lo, z_hi) into out (out points at plane z_lo).
// flags: one device word, bit 0 set when the item pool was too small (the result is then not valid).
// events (may be null): 2 * chunks + 1 events, recorded in front of every kernel and after the last.
int launch_paint_disc(const uint32_t *ridge, int w, int h, int dext, int z_lo, int z_hi, uint32_t rsq_max,
                      int chunk_planes, int items_per_voxel, void *workspace, uint32_t *out, uint32_t *flags,
                      cudaEvent_t *events, cudaStream_t s) {
    const size_t wh = (size_t)w * h;
    const int nxt = (w + 31) / 32;
    const int emax = (int)host_isqrt(rsq_max);
    unsigned long long *cursor = (unsigned long long *)workspace;
    unsigned long long *desc = cursor + 32;
    unsigned long long *pool = desc + (size_t)chunk_planes * h * nxt;
    const unsigned long long pool_cap = (unsigned long long)chunk_planes * wh * (unsigned long long)items_per_voxel;
    const size_t zsm = zfront_smem_bytes(emax);
    cudaFuncSetAttribute(zfront_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zsm);
    cudaFuncSetAttribute(disc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DISC_SMEM);
    const int ntx = (w + DTW - 1) / DTW, nty = (h + DTH - 1) / DTH;
    int launches = 0;
    for (int zc0 = z_lo; zc0 < z_hi; zc0 += chunk_planes) {
        const int zc1 = min(z_hi, zc0 + chunk_planes);
        cudaMemsetAsync(cursor, 0, 8, s);
        if (events) cudaEventRecord(events[launches], s);
        const unsigned nseg = (unsigned)((zc1 - zc0 + ZF_ZB - 1) / ZF_ZB);
        zfront_kernel<<<(unsigned)nxt * (unsigned)h * nseg, ZF_THREADS, zsm, s>>>(ridge, w, h, dext, zc0, zc1, emax, nxt, desc,
                                                                                  pool, pool_cap, cursor, flags);
        if (events) cudaEventRecord(events[launches + 1], s);
        disc_kernel<<<dim3((unsigned)(ntx * nty), (unsigned)(zc1 - zc0)), DISC_THREADS, DISC_SMEM, s>>>(
            desc, pool, w, h, nxt, emax, ntx, out + (size_t)(zc0 - z_lo) * wh);
        launches += 2;
    }
    if (events) cudaEventRecord(events[launches], s);
    return launches;
}

}  // namespace bjt
