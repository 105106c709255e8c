// paint_sep.cu -- separable largest-inscribed-sphere painter (paths 0 and 1), sm_100a.
//
// Replaces the scatter of sc.fiji.localThickness.Local_Thickness_Parallel with an
// exact axis-by-axis propagation: out(p) = max{ R_q : |p-q|^2 <= R_q } over ridge points q.  The
// per-line algorithm is in paint_sep_core.cuh (shared with the CPU model
// oracle/sep_paint_model.cpp that tests it without a GPU).  Cost is O(N * front size) instead of
// the sum of ball volumes.
//
// Layout.  One thread per line, every lane of a warp at the same position of its line, so a warp moves
// through the volume as a plane front.  A warp's 32 lines are an 8 x 4 patch of the plane of lines
// (sep_patch_line in paint_sep_core.cuh): neighbours in both directions have similar active sets, which is
// what a lock-step warp pays for.  Fronts are variable-length lists:
// per voxel one u64 index entry (offset:40 | count:24) into an item pool; a warp allocates the
// items of one step with a single atomicAdd (shuffle prefix sum), so items of neighbouring lines
// are neighbours in the pool.  Forward and backward fronts are kept as two lists; the next stage
// reads both.
//   stage 1  lines along x (thread per row), whole volume            ridge -> L1F, L1B
//   stage 2  lines along y, lanes along x, per x-slab of columns     L1  -> L2F, L2B (slab-local)
//   stage 3  lines along z, lanes along x, same slab                  L2  -> painted r^2 (coalesced)
// Items are (slack, R).  PACKED: u32 = R << 16 | slack when the largest r^2 is below 65536;
// otherwise u64.
//
// Active sets.  A line's active set is long-lived state that is touched at every step, while the
// list traffic streams through L1; to keep the former from being evicted by the latter, PACKED
// kernels hold the first S_SM entries of each buffer in shared memory (48-bit entries in u64,
// interleaved by thread) and only the tail in the thread's local frame.  Lines that outgrow the
// frame (a few, in very large pores) are recorded and redone by a second kernel whose active sets
// live in a global scratch; anything beyond that, or a full pool, raises a flag and the caller
// falls back to the scatter painter -- never silently.
#include "common.cuh"
#include <atomic>
#include "paint_sep_core.cuh"

namespace bjt {

namespace {

// x-slab width of stages 2 and 3: as many columns as keep a slab at or below 2^29 voxels (the item pool of
// stage 2 is sized per slab voxel), at least 512 -- thin z-slabs of a sharded run get the whole width, which
// keeps the y-stage grid (columns x planes lines) large enough to fill the GPU
std::atomic<int> g_xslab_override{0};  // paint_sep_set_xslab_cols: the ranks of a sharded run agree on one width; tests
inline int natural_xslab_cols(int h, int d) {          // not clipped to the width of the volume
    const long long per_col = (long long)h * d;
    long long xc = (1ll << 29) / (per_col > 0 ? per_col : 1);
    xc = xc / 128 * 128;
    if (xc < 512) xc = 512;
    return xc < (1 << 20) ? (int)xc : (1 << 20);
}
inline int xslab_cols(int w, int h, int d) {
    const int ov = g_xslab_override.load();
    const int xc = ov > 0 ? ov : natural_xslab_cols(h, d);
    return xc < w ? xc : w;
}
constexpr unsigned long long OFF_MASK = (1ull << 40) - 1ull;
constexpr int TPB = 128;                // threads per block of the line kernels
// Tuning knobs with their measured defaults; tools/tune_painter.sh rebuilds with other values (-DBJT_S_SM=.. etc.)
#ifndef BJT_S_SM
#define BJT_S_SM 12
#endif
#ifndef BJT_DYN_BLOCKS
#define BJT_DYN_BLOCKS 9                // resident blocks per SM the line kernels of stages 1-2 are compiled for (56 registers)
#endif
#ifndef BJT_STAIR_BLOCKS
#define BJT_STAIR_BLOCKS 12             // the same for the stage-3 kernel (40 registers)
#endif
constexpr int S_SM = BJT_S_SM;          // entries per buffer kept in shared memory (PACKED kernels)
constexpr int CAP_TAIL = 1024 - S_SM;   // entries per buffer in the local frame behind them
constexpr int CAP_WIDE = 128;           // local-frame capacity of the wide (u64 item) kernels
constexpr int FMAX_FAST = 256;
constexpr int CAP_STAIR = 96;         // local-frame capacity of the stage-3 staircase
constexpr int CAP_BIG = 4096, FMAX_BIG = 1024, NSLOTS = 2048;

enum : uint32_t { F_POOL1 = 1u, F_POOL2 = 2u, F_ACTIVE = 4u, F_FRONT = 8u };

// test hook: caps on the fast kernels' capacities (paint_sep_set_limits) so that small volumes can drive
// lines into the redo kernels; the defaults never bind
__constant__ int c_cap_limit = 1 << 30, c_fmax_limit = 1 << 30;

template <bool PACKED> struct ItemIO;
template <> struct ItemIO<true> {
    typedef uint32_t T;
    static __device__ __forceinline__ T pack(SepItem it) { return (it.R << 16) | it.rho; }
    static __device__ __forceinline__ T raw(const T *p, int k) { return __ldg(p + k); }
    static __device__ __forceinline__ SepItem unpack(T v) { SepItem it; it.rho = v & 0xFFFFu; it.R = v >> 16; return it; }
    static __device__ __forceinline__ SepItem get(const T *p, int k) { return unpack(raw(p, k)); }
};
template <> struct ItemIO<false> {
    typedef uint2 T;
    static __device__ __forceinline__ T pack(SepItem it) { return make_uint2(it.rho, it.R); }
    static __device__ __forceinline__ T raw(const T *p, int k) { return __ldg(p + k); }
    static __device__ __forceinline__ SepItem unpack(T v) { SepItem it; it.rho = v.x; it.R = v.y; return it; }
    static __device__ __forceinline__ SepItem get(const T *p, int k) { return unpack(raw(p, k)); }
};
struct PlainIO { static __device__ __forceinline__ SepItem get(const SepItem *p, int k) { return p[k]; } };

// PACKED active-set storage: R < 2^16, rho0 < 2^16, |t| < 2^15 -> one u64 per entry, low word laid out
// like a packed item (R << 16 | rho0), high word t + 32768.
// Entry (half, i): i < S_SM in shared memory at word (half * S_SM + i) * TPB of the thread's column (a warp
// touches 32 consecutive u64), the rest in the local frame.
struct HybridStore {
    typedef unsigned long long Entry;
    uint32_t sm;                           // shared-window address of the thread's column
    unsigned long long loc[2][CAP_TAIL];
    static __device__ __forceinline__ Entry make(uint32_t R, uint32_t rho0, int t) {
        return (unsigned long long)((R << 16) | rho0) | ((unsigned long long)(uint32_t)(t + 32768) << 32);
    }
    static __device__ __forceinline__ uint32_t R(Entry e) { return (uint32_t)e >> 16; }
    static __device__ __forceinline__ uint32_t rho0(Entry e) { return (uint32_t)e & 0xFFFFu; }
    static __device__ __forceinline__ int t(Entry e) { return (int)(uint32_t)(e >> 32) - 32768; }
    __device__ __forceinline__ int cap() const { return min(S_SM + CAP_TAIL, c_cap_limit); }
    __device__ __forceinline__ Entry load(int half, int i) const {
        if (i < S_SM) {
            Entry v;
            asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(sm + (uint32_t)((half * S_SM + i) * TPB * 8)));
            return v;
        }
        return loc[half][i - S_SM];
    }
    __device__ __forceinline__ void store(int half, int i, Entry e) {
        if (i < S_SM) asm volatile("st.shared.u64 [%0], %1;" ::"r"(sm + (uint32_t)((half * S_SM + i) * TPB * 8)), "l"(e) : "memory");
        else loc[half][i - S_SM] = e;
    }
};

// warp-aggregated append of nf items per lane; returns the lane's offset or ~0 on pool overflow.
// A warp takes CHUNK items at a time from the global counter (one same-address atomic per chunk, not
// per step) and hands them out from registers; chunk_base / chunk_left are warp-uniform.
constexpr int CHUNK = 4096;
__device__ __forceinline__ unsigned long long warp_alloc(int nf, unsigned long long *counter, unsigned long long cap,
                                                         uint32_t *flags, uint32_t flagbit,
                                                         unsigned long long &chunk_base, int &chunk_left) {
    const int lane = threadIdx.x & 31;
    if (!__any_sync(0xFFFFFFFFu, nf > 0)) return 0ull;      // nothing emitted by the whole warp: common where a run has no balls
    int incl = nf;
#pragma unroll
    for (int k = 1; k < 32; k <<= 1) {
        const int u = __shfl_up_sync(0xFFFFFFFFu, incl, k);
        if (lane >= k) incl += u;
    }
    const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    if (total == 0) return 0ull;
    if (total > chunk_left) {
        const int take = total > CHUNK ? total : CHUNK;
        unsigned long long base = 0ull;
        if (lane == 0) base = atomicAdd(counter, (unsigned long long)take);
        chunk_base = __shfl_sync(0xFFFFFFFFu, base, 0);
        chunk_left = take;
        if (chunk_base + (unsigned long long)take > cap) {
            if (lane == 0) atomicOr(flags, flagbit);
            chunk_left = 0;
            return ~0ull;
        }
    }
    const unsigned long long off = chunk_base + (unsigned long long)(incl - nf);
    chunk_base += (unsigned long long)total;
    chunk_left -= total;
    return off;
}

// non-binding request for the line holding p (L1): a semantic no-op that shortens a later load
__device__ __forceinline__ void prefetch_l1(const void *p) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(__cvta_generic_to_global(p)));
}

// How far ahead the line kernels look.  The index words of a position are loaded one position ahead (registers);
// the items they point to are requested PF_ITEMS positions ahead (L1): the index word of that position is itself
// requested into L1 at the top of a position and read at its bottom (an L1 hit), so no register is held across the
// position and nothing waits on memory.  The round's ncu captures had 19 % of the
// y-stage sweep's samples and 15 % of the z-stage's on the first use of an item loaded on demand, and 10 % of
// the x-stage's on the ridge value loaded one position ahead (its rows are read 4 bytes per lane).
constexpr int PF_ITEMS = 2;             // positions
constexpr int PF_RIDGE = 16;            // positions (two 32-byte sectors)

struct SepArgs {
    // stage 1 input
    const uint32_t *ridge;
    // stage 2/3 input lists
    const unsigned long long *inF, *inB;
    const void *in_pool;
    // geometry
    int w, h, d;            // whole volume
    int xc0, xw;            // x-slab (stages 2, 3)
    int dir;
    // stage 1/2 output lists
    unsigned long long *out_idx;
    void *out_pool;
    unsigned long long out_cap;
    unsigned long long *out_counter;
    uint32_t poolflag;
    // stage 3 output
    uint32_t *out;
    // bookkeeping
    uint32_t *flags;
    int *redo_list;          // lines whose fast capacities were too small
    int *redo_count;
    unsigned long long *redo_total;     // lines sent to the redo kernels in this run (diagnostics, tests)
    char *scratch;           // stores of the redo kernels
    // stage 3 of a z-slab: staircases the neighbouring slabs exported (arrive at plane 0 / plane d-1)
    SepImports imp;
};

// one line of stage 1 (STAGE == 1: line = row (y,z), position = x) or stage 2 (line = (lx, z), position = y).
// All 32 lanes of the warp must call this together (warp-aggregated allocation); `valid` masks lanes.
template <bool PACKED, int STAGE, int DIR, class Active>
__device__ __forceinline__ bool dyn_line(const SepArgs &a, long long line, bool valid, Active &A, SepItem *fr, int fmax) {
    typedef ItemIO<PACKED> IO;
    const int L = STAGE == 1 ? a.w : a.h;
    long long in_base = 0, in_stride = 0, out_base = 0, out_stride = 0;
    if (STAGE == 1) {
        in_base = line * a.w; in_stride = 1;
        out_base = in_base; out_stride = 1;
    } else {
        const int lx = (int)(line % a.xw);
        const long long z = line / a.xw;
        in_base = z * a.h * (long long)a.w + a.xc0 + lx; in_stride = a.w;
        out_base = z * a.h * (long long)a.xw + lx; out_stride = a.xw;
    }
    A.clear();
    bool ok = true;
    const typename IO::T *ipool = reinterpret_cast<const typename IO::T *>(a.in_pool);
    typename IO::T *opool = reinterpret_cast<typename IO::T *>(a.out_pool);
    unsigned long long chunk_base = 0ull;
    int chunk_left = 0;
    // the inputs of the next position are requested before this position is processed
    const long long step_in = DIR > 0 ? in_stride : -in_stride;
    long long vi = in_base + (long long)(DIR > 0 ? 0 : L - 1) * in_stride;
    uint32_t r_next = 0;
    unsigned long long eF_next = 0ull, eB_next = 0ull;
    if (valid) {
        if (STAGE == 1) r_next = __ldg(a.ridge + vi);
        else { eF_next = __ldg(a.inF + vi); eB_next = __ldg(a.inB + vi); }
    }
    for (int step = 0; step < L; ++step) {
        const int t = DIR > 0 ? step : L - 1 - step;
        int nf = 0;
        if (valid) {
            const uint32_t r = r_next;
            const unsigned long long eF = eF_next, eB = eB_next;
            if (step + 1 < L) {
                vi += step_in;
                if (STAGE == 1) r_next = __ldg(a.ridge + vi);
                else { eF_next = __ldg(a.inF + vi); eB_next = __ldg(a.inB + vi); }
            }
            if (STAGE == 1) {
                if (step + PF_RIDGE < L) prefetch_l1(a.ridge + vi + (long long)(PF_RIDGE - 1) * step_in);
            } else if (step + PF_ITEMS < L) {                // their index words first (no register held across the step)
                prefetch_l1(a.inF + vi + (long long)(PF_ITEMS - 1) * step_in);
                prefetch_l1(a.inB + vi + (long long)(PF_ITEMS - 1) * step_in);
            }
            if (STAGE == 1) {
                if (r | (uint32_t)A.n) {
                    SepItem one; one.rho = r; one.R = r;
                    nf = A.template step<SepItem, PlainIO, false, DIR>(t, &one, r ? 1 : 0, (const SepItem *)nullptr, 0, fr, fmax, ok);
                }
            } else {
                if (eF | eB | (unsigned long long)A.n)
                    nf = A.template step<typename IO::T, IO, true, DIR>(t, ipool + (eF & OFF_MASK), (int)(eF >> 40),
                                                             ipool + (eB & OFF_MASK), (int)(eB >> 40), fr, fmax, ok);
            }
            if (nf > fmax) { ok = false; nf = fmax; }
            if (STAGE != 1 && step + PF_ITEMS < L) {         // the lists of position t + PF_ITEMS, into L1
                const unsigned long long pfF = __ldg(a.inF + vi + (long long)(PF_ITEMS - 1) * step_in);
                const unsigned long long pfB = __ldg(a.inB + vi + (long long)(PF_ITEMS - 1) * step_in);
                if (pfF >> 40) prefetch_l1(ipool + (pfF & OFF_MASK));
                if (pfB >> 40) prefetch_l1(ipool + (pfB & OFF_MASK));
            }
        }
        const unsigned long long off = warp_alloc(nf, a.out_counter, a.out_cap, a.flags, a.poolflag, chunk_base, chunk_left);
        if (valid) {
            unsigned long long entry = 0ull;
            if (nf > 0 && off != ~0ull) {
                for (int k = 0; k < nf; ++k) opool[off + k] = IO::pack(fr[k]);
                entry = off | ((unsigned long long)nf << 40);
            }
            a.out_idx[out_base + (long long)t * out_stride] = entry;
        }
    }
    return ok;
}

// one line of stage 3: line = (lx, y), position = z
template <bool PACKED, class Stair>
__device__ __forceinline__ bool stair_line(const SepArgs &a, long long line, Stair &S) {
    typedef ItemIO<PACKED> IO;
    const int lx = (int)(line % a.xw);
    const long long y = line / a.xw;
    const long long in_base = y * a.xw + lx, in_stride = (long long)a.h * a.xw;
    const long long out_base = y * a.w + a.xc0 + lx, out_stride = (long long)a.h * a.w;
    const typename IO::T *ipool = reinterpret_cast<const typename IO::T *>(a.in_pool);
    S.clear();
    bool ok = true;
    const long long step_in = a.dir > 0 ? in_stride : -in_stride;
    long long vi = in_base + (long long)(a.dir > 0 ? 0 : a.d - 1) * in_stride;
    unsigned long long eF_next = __ldg(a.inF + vi), eB_next = __ldg(a.inB + vi);
    // the backward sweep keeps the larger of its value and what the forward sweep left: that old value is read one
    // position ahead, so the comparison at the end of a position does not wait for it (5.6 % of the backward sweep's
    // samples sat on that comparison).  The line's output voxels belong to this thread alone.
    uint32_t o_next = 0;
    if (a.dir < 0) o_next = a.out[out_base + (long long)(a.d - 1) * out_stride];
    for (int step = 0; step < a.d; ++step) {
        const int z = a.dir > 0 ? step : a.d - 1 - step;
        const int u = a.dir > 0 ? z : -z;
        const unsigned long long eF = eF_next, eB = eB_next;
        if (step + 1 < a.d) { vi += step_in; eF_next = __ldg(a.inF + vi); eB_next = __ldg(a.inB + vi); }
        const uint32_t o_old = o_next;
        if (a.dir < 0 && step + 1 < a.d) o_next = a.out[out_base + (long long)(z - 1) * out_stride];
        if (step + PF_ITEMS < a.d) {                        // index words PF_ITEMS positions ahead: no register held across the step
            prefetch_l1(a.inF + vi + (long long)(PF_ITEMS - 1) * step_in);
            prefetch_l1(a.inB + vi + (long long)(PF_ITEMS - 1) * step_in);
        }
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            const unsigned long long e = which ? eB : eF;
            const typename IO::T *p = ipool + (e & OFF_MASK);
            const int cnt = (int)(e >> 40);
            int hint = 0;                                   // tags descend within a list: each search resumes at the last
            if (cnt > 0) {
                typename IO::T v = IO::raw(p, 0);
                for (int k = 0; k < cnt; ++k) {             // the next item is on its way while this one is inserted
                    const SepItem it = IO::unpack(v);
                    if (k + 1 < cnt) v = IO::raw(p, k + 1);
                    ok &= S.arrive(u, it.rho, it.R, hint);
                }
            }
        }
        if (step + PF_ITEMS < a.d) {                        // the lists of plane z +- PF_ITEMS, into L1
            const unsigned long long pfF = __ldg(a.inF + vi + (long long)(PF_ITEMS - 1) * step_in);
            const unsigned long long pfB = __ldg(a.inB + vi + (long long)(PF_ITEMS - 1) * step_in);
            if (pfF >> 40) prefetch_l1(ipool + (pfF & OFF_MASK));
            if (pfB >> 40) prefetch_l1(ipool + (pfB & OFF_MASK));
        }
        if (z == 0) {
            for (int src = 0; src < a.imp.n_lo; ++src) {
                const int cnt = (int)a.imp.lo_cnt[src][in_base];
                const uint2 *p = a.imp.lo_items[src] + in_base * a.imp.kcap;
                int hint = 0;
                for (int k = 0; k < cnt; ++k) { const uint2 it = p[k]; ok &= S.arrive(u, it.x, it.y, hint); }
            }
        }
        if (z == a.d - 1) {
            for (int src = 0; src < a.imp.n_hi; ++src) {
                const int cnt = (int)a.imp.hi_cnt[src][in_base];
                const uint2 *p = a.imp.hi_items[src] + in_base * a.imp.kcap;
                int hint = 0;
                for (int k = 0; k < cnt; ++k) { const uint2 it = p[k]; ok &= S.arrive(u, it.x, it.y, hint); }
            }
        }
        const uint32_t r = S.query(u);
        uint32_t *o = a.out + out_base + (long long)z * out_stride;
        if (a.dir > 0) *o = r;
        else if (r > o_old) *o = r;
    }
    return ok;
}

__device__ __forceinline__ void record_redo(const SepArgs &a, long long line) {
    const int slot = atomicAdd(a.redo_count, 1);
    a.redo_list[slot] = (int)line;
}

template <bool PACKED> struct FastActive;
template <> struct FastActive<true> { typedef SepActive<HybridStore> T; };
struct WideStore : SepLocalStore<CAP_WIDE> {
    __device__ __forceinline__ int cap() const { return min(CAP_WIDE, c_cap_limit); }
};
struct StairStore : SepStairLocal<CAP_STAIR> {
    __device__ __forceinline__ int cap() const { return min(CAP_STAIR, c_cap_limit); }
};
template <> struct FastActive<false> { typedef SepActive<WideStore> T; };

template <bool PACKED, int STAGE, int DIR>
__global__ void __launch_bounds__(TPB, BJT_DYN_BLOCKS) sep_dyn_kernel(SepArgs a) {
    __shared__ unsigned long long sm[PACKED ? 2 * S_SM * TPB : 1];
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long pl = STAGE == 1 ? sep_patch_line(tid, a.h, a.d) : sep_patch_line(tid, a.xw, a.d);
    const bool valid = pl >= 0;                      // whole warps stay in: dyn_line allocates as a warp
    const long long line = valid ? pl : 0;
    typename FastActive<PACKED>::T A;
    if constexpr (PACKED) A.st.sm = (uint32_t)__cvta_generic_to_shared(sm + threadIdx.x);
    SepItem fr[FMAX_FAST];
    const bool ok = dyn_line<PACKED, STAGE, DIR>(a, line, valid, A, fr, min(FMAX_FAST, c_fmax_limit));
    if (valid && !ok) record_redo(a, line);
}

// per-slot scratch of the redo kernels: 2 * CAP_BIG entries, then FMAX_BIG front items
constexpr size_t SLOT_BYTES = (size_t)2 * CAP_BIG * sizeof(SepEntry) + (size_t)FMAX_BIG * sizeof(SepItem);

// redo: NSLOTS threads, each takes failed lines slot, slot + NSLOTS, ...
template <bool PACKED, int STAGE, int DIR>
__global__ void __launch_bounds__(32) sep_dyn_redo_kernel(SepArgs a) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    const int count = *a.redo_count;
    if (count == 0) return;
    if (slot == 0) atomicAdd(a.redo_total, (unsigned long long)count);
    SepActive<SepFlatStore> A;
    char *mine = a.scratch + (size_t)slot * SLOT_BYTES;
    A.st.base = reinterpret_cast<SepEntry *>(mine); A.st.capacity = CAP_BIG;
    SepItem *fr = reinterpret_cast<SepItem *>(mine + (size_t)2 * CAP_BIG * sizeof(SepEntry));
    for (int base = 0; base < count; base += NSLOTS) {      // warp-uniform trip count
        const int idx = base + slot;
        const bool valid = idx < count;
        const long long line = valid ? a.redo_list[idx] : 0;
        const bool ok = dyn_line<PACKED, STAGE, DIR>(a, line, valid, A, fr, FMAX_BIG);
        if (valid && !ok) atomicOr(a.flags, (uint32_t)F_ACTIVE);
    }
}

template <bool PACKED>
__global__ void __launch_bounds__(TPB, BJT_STAIR_BLOCKS) sep_stair_kernel(SepArgs a) {
    const long long line = sep_patch_line((long long)blockIdx.x * blockDim.x + threadIdx.x, a.xw, a.h);
    if (line < 0) return;
    SepStair<StairStore> S;
    if (!stair_line<PACKED>(a, line, S)) record_redo(a, line);
}

template <bool PACKED>
__global__ void __launch_bounds__(32) sep_stair_redo_kernel(SepArgs a) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    const int count = *a.redo_count;
    if (slot == 0 && count) atomicAdd(a.redo_total, (unsigned long long)count);
    SepStair<SepStairFlat> S;
    S.st.R = reinterpret_cast<uint32_t *>(a.scratch + (size_t)slot * SLOT_BYTES);
    S.st.E = reinterpret_cast<int32_t *>(S.st.R + 2 * CAP_BIG);
    S.st.capacity = 2 * CAP_BIG;
    for (int idx = slot; idx < count; idx += NSLOTS)
        if (!stair_line<PACKED>(a, a.redo_list[idx], S)) atomicOr(a.flags, (uint32_t)F_ACTIVE);
}

// boundary staircase of a z-slab: what the items of the planes next to a slab face still cover beyond it.
// One thread per column; sweep towards the face over the last hh planes, then write the steps that reach the
// receiver's first plane (coordinate `target` in sweep direction) as (reach beyond it, tag), tags descending.
template <bool PACKED>
__global__ void __launch_bounds__(TPB) sep_export_kernel(SepArgs a, int side, int gap, int hh, uint32_t *cnt, uint2 *items,
                                                          int kcap) {
    typedef ItemIO<PACKED> IO;
    const long long line = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= (long long)a.xw * a.h) return;
    const long long in_stride = (long long)a.h * a.xw;
    const typename IO::T *ipool = reinterpret_cast<const typename IO::T *>(a.in_pool);
    SepStair<StairStore> S;
    S.clear();
    bool ok = true;
    const int nplanes = hh < a.d ? hh : a.d;
    const int target = (side > 0 ? a.d - 1 : 0) + gap;
    for (int step = 0; step < nplanes; ++step) {
        const int z = side > 0 ? a.d - nplanes + step : nplanes - 1 - step;
        const int u = side > 0 ? z : -z;
        const long long vi = line + (long long)z * in_stride;
        const unsigned long long eF = __ldg(a.inF + vi), eB = __ldg(a.inB + vi);
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            const unsigned long long e = which ? eB : eF;
            const typename IO::T *p = ipool + (e & OFF_MASK);
            const int c = (int)(e >> 40);
            int hint = 0;
            for (int k = 0; k < c; ++k) {
                const SepItem it = IO::get(p, k);
                if (u + (int)it.rho < target) continue;                     // does not reach the receiver
                ok &= S.arrive(u, it.rho, it.R, hint);
            }
        }
    }
    int nout = 0;
    for (int i = 0; i < S.n; ++i) {
        const int E = S.st.getE(i);
        if (E < target) continue;
        if (nout < kcap) items[line * kcap + nout] = make_uint2((uint32_t)(E - target), S.st.getR(i));
        ++nout;
    }
    if (nout > kcap) { ok = false; nout = kcap; }
    cnt[line] = (uint32_t)nout;
    if (!ok) atomicOr(a.flags, (uint32_t)F_FRONT);
}

struct Plan {
    size_t n, nslab, maxlines;
    size_t off_idx1f, off_idx1b, off_idx2f, off_idx2b, off_pool1, off_pool2, off_counters, off_redo, off_scratch, total;
    unsigned long long cap1, cap2;
};

Plan make_plan(int w, int h, int d, int wide, int scale) {
    Plan p;
    p.n = (size_t)w * h * d;
    const int xw = xslab_cols(w, h, d);
    p.nslab = (size_t)xw * h * d;
    const size_t l1 = (size_t)h * d, l2 = (size_t)xw * d, l3 = (size_t)xw * h;
    p.maxlines = l1 > l2 ? (l1 > l3 ? l1 : l3) : (l2 > l3 ? l2 : l3);
    const size_t isz = wide ? 8 : 4;
    // items; every warp may leave up to one CHUNK unused per sweep
    p.cap1 = (unsigned long long)p.n * (unsigned)(4 * scale) + (unsigned long long)sep_patch_warps(h, d) * CHUNK * 2 + 65536;
    p.cap2 = (unsigned long long)p.nslab * (unsigned)(12 * scale) + (unsigned long long)sep_patch_warps(xw, d) * CHUNK * 2 + 65536;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 255) & ~(size_t)255; return r; };
    p.off_counters = take(256);
    p.off_idx1f = take(p.n * 8); p.off_idx1b = take(p.n * 8);
    p.off_idx2f = take(p.nslab * 8); p.off_idx2b = take(p.nslab * 8);
    p.off_pool1 = take(p.cap1 * isz); p.off_pool2 = take(p.cap2 * isz);
    p.off_redo = take(p.maxlines * sizeof(int));
    p.off_scratch = take((size_t)NSLOTS * SLOT_BYTES);
    p.total = o;
    return p;
}

struct Bufs {
    unsigned long long *counters, *idx1f, *idx1b, *idx2f, *idx2b;
    int *redo_count;
    void *pool1, *pool2;
    SepArgs base;
};

Bufs make_bufs(int w, int h, int d, const Plan &p, char *ws, uint32_t *flags) {
    Bufs b;
    b.counters = reinterpret_cast<unsigned long long *>(ws + p.off_counters);
    b.redo_count = reinterpret_cast<int *>(b.counters + 4);
    b.idx1f = reinterpret_cast<unsigned long long *>(ws + p.off_idx1f);
    b.idx1b = reinterpret_cast<unsigned long long *>(ws + p.off_idx1b);
    b.idx2f = reinterpret_cast<unsigned long long *>(ws + p.off_idx2f);
    b.idx2b = reinterpret_cast<unsigned long long *>(ws + p.off_idx2b);
    b.pool1 = ws + p.off_pool1; b.pool2 = ws + p.off_pool2;
    b.base = SepArgs{};
    b.base.w = w; b.base.h = h; b.base.d = d; b.base.flags = flags;
    b.base.redo_list = reinterpret_cast<int *>(ws + p.off_redo); b.base.redo_count = b.redo_count;
    b.base.redo_total = b.counters + 5;
    b.base.scratch = ws + p.off_scratch;
    return b;
}

// stage 1: both directions share pool 1
template <bool PACKED>
int run_stage1(const uint32_t *ridge, int w, int h, int d, const Plan &p, char *ws, uint32_t *flags, cudaStream_t s) {
    Bufs b = make_bufs(w, h, d, p, ws, flags);
    int launches = 0;
    cudaMemsetAsync(b.counters, 0, 256, s);
    for (int dir = +1; dir >= -1; dir -= 2) {
        SepArgs a = b.base;
        a.ridge = ridge; a.xc0 = 0; a.xw = w; a.dir = dir;
        a.out_idx = dir > 0 ? b.idx1f : b.idx1b; a.out_pool = b.pool1; a.out_cap = p.cap1; a.out_counter = b.counters;
        a.poolflag = F_POOL1;
        cudaMemsetAsync(b.redo_count, 0, 4, s);
        const unsigned grid = (unsigned)((sep_patch_threads(h, d) + TPB - 1) / TPB);
        if (dir > 0) {
            sep_dyn_kernel<PACKED, 1, +1><<<grid, TPB, 0, s>>>(a);
            sep_dyn_redo_kernel<PACKED, 1, +1><<<NSLOTS / 32, 32, 0, s>>>(a);
        } else {
            sep_dyn_kernel<PACKED, 1, -1><<<grid, TPB, 0, s>>>(a);
            sep_dyn_redo_kernel<PACKED, 1, -1><<<NSLOTS / 32, 32, 0, s>>>(a);
        }
        launches += 2;
    }
    return launches;
}

// stage 2 of one x-slab: lists along y into pool 2 (reused by every slab)
template <bool PACKED>
int run_stage2(int w, int h, int d, const Plan &p, char *ws, int xc0, uint32_t *flags, cudaStream_t s) {
    Bufs b = make_bufs(w, h, d, p, ws, flags);
    int launches = 0;
    const int XC = xslab_cols(w, h, d);
    const int xw = (w - xc0) < XC ? (w - xc0) : XC;
    cudaMemsetAsync(b.counters + 1, 0, 8, s);
    for (int dir = +1; dir >= -1; dir -= 2) {
        SepArgs a = b.base;
        a.inF = b.idx1f; a.inB = b.idx1b; a.in_pool = b.pool1; a.xc0 = xc0; a.xw = xw; a.dir = dir;
        a.out_idx = dir > 0 ? b.idx2f : b.idx2b; a.out_pool = b.pool2; a.out_cap = p.cap2; a.out_counter = b.counters + 1;
        a.poolflag = F_POOL2;
        cudaMemsetAsync(b.redo_count, 0, 4, s);
        const unsigned grid = (unsigned)((sep_patch_threads(xw, d) + TPB - 1) / TPB);
        if (dir > 0) {
            sep_dyn_kernel<PACKED, 2, +1><<<grid, TPB, 0, s>>>(a);
            sep_dyn_redo_kernel<PACKED, 2, +1><<<NSLOTS / 32, 32, 0, s>>>(a);
        } else {
            sep_dyn_kernel<PACKED, 2, -1><<<grid, TPB, 0, s>>>(a);
            sep_dyn_redo_kernel<PACKED, 2, -1><<<NSLOTS / 32, 32, 0, s>>>(a);
        }
        launches += 2;
    }
    return launches;
}

// stage 3 of one x-slab: z sweeps, painted r^2 out
template <bool PACKED>
int run_stage3(int w, int h, int d, const Plan &p, char *ws, int xc0, const SepImports *imp, uint32_t *rsq, uint32_t *flags,
               cudaStream_t s) {
    Bufs b = make_bufs(w, h, d, p, ws, flags);
    int launches = 0;
    const int XC = xslab_cols(w, h, d);
    const int xw = (w - xc0) < XC ? (w - xc0) : XC;
    for (int dir = +1; dir >= -1; dir -= 2) {
        SepArgs a = b.base;
        a.inF = b.idx2f; a.inB = b.idx2b; a.in_pool = b.pool2; a.xc0 = xc0; a.xw = xw; a.dir = dir;
        a.out = rsq;
        if (imp) a.imp = *imp;
        cudaMemsetAsync(b.redo_count, 0, 4, s);
        sep_stair_kernel<PACKED><<<(unsigned)((sep_patch_threads(xw, h) + TPB - 1) / TPB), TPB, 0, s>>>(a);
        sep_stair_redo_kernel<PACKED><<<NSLOTS / 32, 32, 0, s>>>(a);
        launches += 2;
    }
    return launches;
}

template <bool PACKED>
int run_export(int w, int h, int d, const Plan &p, char *ws, int xc0, int side, int gap, int hh, uint32_t *cnt, uint2 *items,
               int kcap, uint32_t *flags, cudaStream_t s) {
    Bufs b = make_bufs(w, h, d, p, ws, flags);
    const int XC = xslab_cols(w, h, d);
    const int xw = (w - xc0) < XC ? (w - xc0) : XC;
    SepArgs a = b.base;
    a.inF = b.idx2f; a.inB = b.idx2b; a.in_pool = b.pool2; a.xc0 = xc0; a.xw = xw; a.dir = side;
    const long long nlines = (long long)xw * h;
    sep_export_kernel<PACKED><<<(unsigned)((nlines + TPB - 1) / TPB), TPB, 0, s>>>(a, side, gap, hh, cnt, items, kcap);
    return 1;
}

template <bool PACKED>
int run_sep(const uint32_t *ridge, int w, int h, int d, const Plan &p, char *ws, uint32_t *rsq, uint32_t *flags,
            cudaStream_t s) {
    int launches = run_stage1<PACKED>(ridge, w, h, d, p, ws, flags, s);
    const int XC = xslab_cols(w, h, d);
    for (int xc0 = 0; xc0 < w; xc0 += XC) {
        launches += run_stage2<PACKED>(w, h, d, p, ws, xc0, flags, s);
        launches += run_stage3<PACKED>(w, h, d, p, ws, xc0, nullptr, rsq, flags, s);
    }
    return launches;
}

}  // namespace

void paint_sep_set_xslab_cols(int cols) { g_xslab_override = cols; }
int paint_sep_natural_xslab_cols(int h, int d) { return natural_xslab_cols(h, d); }

void paint_sep_set_limits(int cap, int fmax) {
    const int c = cap > 0 ? cap : 1 << 30, f = fmax > 0 ? fmax : 1 << 30;
    cudaMemcpyToSymbol(c_cap_limit, &c, sizeof(int));
    cudaMemcpyToSymbol(c_fmax_limit, &f, sizeof(int));
}

// lines the last run on this workspace sent to the redo kernels (synchronous read)
unsigned long long paint_sep_redo_lines(const void *workspace) {
    unsigned long long v = 0;
    cudaMemcpy(&v, reinterpret_cast<const unsigned long long *>(workspace) + 5, 8, cudaMemcpyDeviceToHost);
    return v;
}

// variant: bit 0 = wide (u64) items, bit 1 = doubled pools
size_t paint_sep_workspace_bytes(int w, int h, int d, int variant) {
    return make_plan(w, h, d, variant & 1, (variant & 2) ? 2 : 1).total;
}

int paint_sep_nxslabs(int w, int h, int d) { const int XC = xslab_cols(w, h, d); return (w + XC - 1) / XC; }
int paint_sep_xslab_width(int w, int h, int d, int xs) {
    const int XC = xslab_cols(w, h, d), xc0 = xs * XC;
    return (w - xc0) < XC ? (w - xc0) : XC;
}

int paint_sep_stage1(const uint32_t *ridge, int w, int h, int d, int variant, void *workspace, uint32_t *flags, cudaStream_t s) {
    const Plan p = make_plan(w, h, d, variant & 1, (variant & 2) ? 2 : 1);
    char *ws = reinterpret_cast<char *>(workspace);
    return (variant & 1) ? run_stage1<false>(ridge, w, h, d, p, ws, flags, s) : run_stage1<true>(ridge, w, h, d, p, ws, flags, s);
}
int paint_sep_stage2(int w, int h, int d, int variant, void *workspace, int xs, uint32_t *flags, cudaStream_t s) {
    const Plan p = make_plan(w, h, d, variant & 1, (variant & 2) ? 2 : 1);
    char *ws = reinterpret_cast<char *>(workspace);
    return (variant & 1) ? run_stage2<false>(w, h, d, p, ws, xs * xslab_cols(w, h, d), flags, s) : run_stage2<true>(w, h, d, p, ws, xs * xslab_cols(w, h, d), flags, s);
}
int paint_sep_export(int w, int h, int d, int variant, void *workspace, int xs, int side, int gap, int hh, uint32_t *cnt,
                     uint2 *items, int kcap, uint32_t *flags, cudaStream_t s) {
    const Plan p = make_plan(w, h, d, variant & 1, (variant & 2) ? 2 : 1);
    char *ws = reinterpret_cast<char *>(workspace);
    return (variant & 1) ? run_export<false>(w, h, d, p, ws, xs * xslab_cols(w, h, d), side, gap, hh, cnt, items, kcap, flags, s)
                         : run_export<true>(w, h, d, p, ws, xs * xslab_cols(w, h, d), side, gap, hh, cnt, items, kcap, flags, s);
}
int paint_sep_stage3(int w, int h, int d, int variant, void *workspace, int xs, const SepImports *imp, uint32_t *rsq,
                     uint32_t *flags, cudaStream_t s) {
    const Plan p = make_plan(w, h, d, variant & 1, (variant & 2) ? 2 : 1);
    char *ws = reinterpret_cast<char *>(workspace);
    return (variant & 1) ? run_stage3<false>(w, h, d, p, ws, xs * xslab_cols(w, h, d), imp, rsq, flags, s)
                         : run_stage3<true>(w, h, d, p, ws, xs * xslab_cols(w, h, d), imp, rsq, flags, s);
}

int launch_paint_sep(const uint32_t *ridge, int w, int h, int d, int variant, void *workspace, uint32_t *rsq,
                     uint32_t *overflow_flag, cudaStream_t s) {
    const int wide = variant & 1;
    const Plan p = make_plan(w, h, d, wide, (variant & 2) ? 2 : 1);
    char *ws = reinterpret_cast<char *>(workspace);
    return wide ? run_sep<false>(ridge, w, h, d, p, ws, rsq, overflow_flag, s)
                : run_sep<true>(ridge, w, h, d, p, ws, rsq, overflow_flag, s);
}

}  // namespace bjt
