/*
 * cuda_runtime.h -- TEST ONLY: a small CPU emulation of the CUDA constructs this repository's kernels use, so the
 * kernels' own source (bonej2_b200/csrc/*.cu, rewritten only at their <<< >>> launches by tests/cuda_emu/emu_build.py)
 * can be compiled with g++ and executed without a GPU.  It exists to check kernel LOGIC (index arithmetic, staging
 * loops, thread -> line maps, warp-collective protocols, list plumbing) against the oracle in the CPU test run; it
 * says nothing about speed, memory-model subtleties or real hardware scheduling.  Not part of the product: the
 * product library is built by nvcc from the unmodified sources.
 *
 * Model: the blocks of a grid run one after another; the threads of a block are fibers (ucontext) on one OS thread,
 * scheduled round-robin; a fiber runs until it reaches a block barrier or a warp collective, where it yields until
 * all live threads of the block / warp have arrived.  __shared__ variables are function-local statics (one block at a
 * time), dynamic shared memory is one static buffer.  Atomics are plain operations (one OS thread).
 */
#pragma once
#define BJT_CUDA_EMU 1
#define BJT_SINGLE_HOST_THREAD 1   /* api.cu: fibers and function-local statics are not thread-safe */

#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <vector>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __constant__
#define __launch_bounds__(...)

struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct emu_uint3 { unsigned x, y, z; };

typedef int cudaError_t;
enum { cudaSuccess = 0 };
typedef void *cudaStream_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
// ---- the runtime calls of the host driver (api.cu): one "device", memory = the heap, streams and events = bookkeeping ----
#include <time.h>
enum { cudaStreamNonBlocking = 1 };
struct emu_event { double ms; };
typedef emu_event *cudaEvent_t;
struct cudaDeviceProp { char name[256]; int multiProcessorCount, major, minor; };
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    memset(p, 0, sizeof *p);
    strcpy(p->name, "CPU emulation (tests/cuda_emu)");
    p->multiProcessorCount = 1; p->major = 0; p->minor = 0;
    return cudaSuccess;
}
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) { *v = 1; return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char *cudaGetErrorString(cudaError_t) { return "emulated CUDA runtime"; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)(uintptr_t)1; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emu_event{0.0}; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
    timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    e->ms = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    return cudaSuccess;
}
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->ms - a->ms); return cudaSuccess; }
template <class T> inline cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)aligned_alloc(256, (n + 255) / 256 * 256); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
template <class T> inline cudaError_t cudaMallocHost(T **p, size_t n) { *p = (T *)aligned_alloc(256, (n + 255) / 256 * 256); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemGetInfo(size_t *free_b, size_t *total_b) { *free_b = (size_t)64 << 30; *total_b = (size_t)64 << 30; return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(dst, src, n); return cudaSuccess; }

inline cudaError_t cudaMemcpy2DAsync(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, cudaMemcpyKind,
                                     cudaStream_t) {
    for (size_t i = 0; i < height; ++i) memcpy((char *)dst + i * dpitch, (const char *)src + i * spitch, width);
    return cudaSuccess;
}
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) { memset(h, 0, sizeof *h); memcpy(h->reserved, &p, sizeof p); return cudaSuccess; }
inline cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, h.reserved, sizeof *p); return cudaSuccess; }
inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }
inline cudaError_t cudaDeviceCanAccessPeer(int *can, int, int) { *can = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }

#define cudaFuncSetAttribute(...) 0
#define cudaMemsetAsync(p, v, n, s) (memset((p), (v), (n)), 0)
#define cudaMemcpyToSymbol(sym, src, n) (memcpy(&(sym), (src), (n)), 0)
#define cudaMemcpy(dst, src, n, kind) (memcpy((dst), (src), (n)), 0)

namespace emu {

constexpr size_t STACK_BYTES = 512 * 1024;

// One rendezvous per participant mask: the lanes named by the mask of a *_sync call (those of them that are still
// alive) meet, each leaves a value, all pick up the snapshot.  Full-mask calls use one slot, sub-warp groups
// (__match_any_sync peers shuffling among themselves) one slot per mask.
struct Rendezvous {
    unsigned mask = 0, arrived_bits = 0, gen = 0;
    unsigned long long vals[32], snap[32];
};
struct Warp {
    std::vector<Rendezvous> slots;
    unsigned live_bits = 0;
    Rendezvous &slot(unsigned mask) {
        for (auto &r : slots) if (r.mask == mask) return r;
        slots.emplace_back();
        slots.back().mask = mask;
        memset(slots.back().vals, 0, sizeof slots.back().vals);
        return slots.back();
    }
};

struct Fiber {
    ucontext_t ctx;
    char *stack = nullptr;
    bool done = false;
    emu_uint3 tidx;
    int lin = 0;
};

struct Block {
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    int alive = 0, arrived = 0;
    unsigned gen = 0;
    int cur = 0;
    ucontext_t main_ctx;
    emu_uint3 bidx;
    dim3 bdim, gdim;
    std::function<void()> body;
};

inline Block *&blk() { static Block *b = nullptr; return b; }
inline Fiber &cur() { return blk()->fibers[blk()->cur]; }
inline void yield() { swapcontext(&cur().ctx, &blk()->main_ctx); }

inline void release_if_complete(Warp &w, Rendezvous &r) {
    const unsigned need = r.mask & w.live_bits;
    if (r.arrived_bits != 0 && (r.arrived_bits & need) == need) {
        memcpy(r.snap, r.vals, sizeof r.snap);
        memset(r.vals, 0, sizeof r.vals);
        r.arrived_bits = 0;
        ++r.gen;
    }
}
inline void block_release_if_complete(Block &b) {
    if (b.alive > 0 && b.arrived == b.alive) { b.arrived = 0; ++b.gen; }
}

inline void trampoline() {
    Block *b = blk();
    b->body();
    Fiber &f = b->fibers[b->cur];
    f.done = true;
    // an exited thread no longer counts at barriers and collectives (CUDA semantics for exited threads)
    Warp &w = b->warps[f.lin >> 5];
    w.live_bits &= ~(1u << (f.lin & 31));
    for (auto &r : w.slots) release_if_complete(w, r);       // whoever waited for this lane stops waiting
    --b->alive;
    block_release_if_complete(*b);
    swapcontext(&f.ctx, &b->main_ctx);
}

inline void sync_block() {
    Block &b = *blk();
    const unsigned g = b.gen;
    ++b.arrived;
    block_release_if_complete(b);
    while (b.gen == g) yield();
}

// every live lane named by `mask` contributes one 64-bit value; returns the group's values (other lanes: 0).
// The snapshot stays valid until the same group meets again, which needs this lane too.
inline const unsigned long long *warp_gather(unsigned mask, unsigned long long v) {
    Block &b = *blk();
    const int lin = cur().lin;
    Warp &w = b.warps[lin >> 5];
    const size_t idx = (size_t)(&w.slot(mask) - w.slots.data());          // slots may move when another mask appears
    const unsigned g = w.slots[idx].gen;
    w.slots[idx].vals[lin & 31] = v;
    w.slots[idx].arrived_bits |= 1u << (lin & 31);
    release_if_complete(w, w.slots[idx]);
    while (w.slots[idx].gen == g) yield();
    return w.slots[idx].snap;
}
inline unsigned live_lanes() { return blk()->warps[cur().lin >> 5].live_bits; }

template <class F>
void launch(dim3 grid, dim3 block, F &&body) {
    const int nthreads = (int)(block.x * block.y * block.z);
    if (nthreads % 32 != 0) { fprintf(stderr, "emu: block size %d is not a multiple of 32\n", nthreads); abort(); }
    Block b;
    b.bdim = block; b.gdim = grid;
    b.body = body;
    b.fibers.resize(nthreads);
    b.warps.resize(nthreads / 32);
    for (auto &f : b.fibers) f.stack = (char *)malloc(STACK_BYTES);
    blk() = &b;
    for (unsigned bz = 0; bz < grid.z; ++bz) for (unsigned by = 0; by < grid.y; ++by) for (unsigned bx = 0; bx < grid.x; ++bx) {
        b.bidx = emu_uint3{bx, by, bz};
        b.alive = nthreads; b.arrived = 0;
        for (auto &w : b.warps) { w.live_bits = 0xFFFFFFFFu; w.slots.clear(); }
        for (int i = 0; i < nthreads; ++i) {
            Fiber &f = b.fibers[i];
            f.done = false; f.lin = i;
            f.tidx = emu_uint3{(unsigned)i % block.x, ((unsigned)i / block.x) % block.y, (unsigned)i / (block.x * block.y)};
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = STACK_BYTES; f.ctx.uc_link = nullptr;
            makecontext(&f.ctx, (void (*)())trampoline, 0);
        }
        int remaining = nthreads;
        long idle_rounds = 0;
        while (remaining > 0) {
            int progressed = 0;
            for (int i = 0; i < nthreads; ++i) {
                if (b.fibers[i].done) continue;
                b.cur = i;
                swapcontext(&b.main_ctx, &b.fibers[i].ctx);
                if (b.fibers[i].done) { --remaining; ++progressed; }
            }
            // a deadlock (a barrier some live thread never reaches) would spin forever: give up loudly
            idle_rounds = progressed ? 0 : idle_rounds + 1;
            if (idle_rounds > 50000000L) { fprintf(stderr, "emu: no thread finished for a very long time (barrier deadlock?)\n"); abort(); }
        }
    }
    for (auto &f : b.fibers) free(f.stack);
    blk() = nullptr;
}

}  // namespace emu

#define threadIdx (emu::cur().tidx)
#define blockIdx (emu::blk()->bidx)
#define blockDim (emu::blk()->bdim)
#define gridDim (emu::blk()->gdim)

inline void __syncthreads() { emu::sync_block(); }

template <class T> inline T __ldg(const T *p) { return *p; }

// ---- warp collectives (full masks only, as the kernels use them) ----
inline int emu_lane() { return emu::cur().lin & 31; }
template <class T> inline unsigned long long emu_bits(T v) { unsigned long long b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T emu_unbits(unsigned long long b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

template <class T> inline T __shfl_sync(unsigned mask, T v, int src) {
    const unsigned long long *s = emu::warp_gather(mask, emu_bits(v));
    return emu_unbits<T>(s[src & 31]);
}
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned delta) {
    const int lane = emu_lane();
    const unsigned long long *s = emu::warp_gather(mask, emu_bits(v));
    return lane >= (int)delta ? emu_unbits<T>(s[lane - (int)delta]) : v;
}
template <class T> inline T __shfl_down_sync(unsigned mask, T v, unsigned delta) {
    const int lane = emu_lane();
    const unsigned long long *s = emu::warp_gather(mask, emu_bits(v));
    return lane + (int)delta < 32 ? emu_unbits<T>(s[lane + (int)delta]) : v;
}
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask) {
    const int lane = emu_lane();
    const unsigned long long *s = emu::warp_gather(mask, emu_bits(v));
    return emu_unbits<T>(s[(lane ^ lane_mask) & 31]);
}
inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
inline unsigned __ballot_sync(unsigned mask, int pred) {
    const unsigned long long *s = emu::warp_gather(mask, pred ? 1ull : 0ull);
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) m |= s[l] ? (1u << l) : 0u;
    return m;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0u; }
inline void __syncwarp(unsigned mask = 0xFFFFFFFFu) { (void)emu::warp_gather(mask, 0ull); }
inline float __fsqrt_rn(float v) { return sqrtf(v); }
inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
    const unsigned long long *s = emu::warp_gather(mask, v);
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) m = (unsigned)s[l] > m ? (unsigned)s[l] : m;
    return m;
}
inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
    const unsigned long long *s = emu::warp_gather(mask, v);
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) m += (unsigned)s[l];
    return m;
}

// lanes that have not left the kernel: equals the hardware's active mask where lanes only diverge by leaving (the
// grid-stride loops of segstats.cu); NOT a model of arbitrary divergence
inline unsigned __activemask() { return emu::live_lanes(); }
// lanes of `mask` holding the same value as the caller (marks a present lane with bit 32 so that a key of 0 counts)
template <class T> inline unsigned __match_any_sync(unsigned mask, T v) {
    const int lane = emu_lane();
    const unsigned long long mine = (emu_bits(v) & 0xFFFFFFFFull) | (1ull << 32);
    static_assert(sizeof(T) <= 4, "emulated __match_any_sync: 32-bit keys");
    const unsigned long long *s = emu::warp_gather(mask, mine);
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) if (s[l] == mine) m |= 1u << l;
    (void)lane;
    return m;
}

// ---- atomics: one OS thread, so plain read-modify-write ----
template <class T> inline T atomicAdd(T *p, T v) { const T o = *p; *p = o + v; return o; }
template <class T> inline T atomicMax(T *p, T v) { const T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicOr(T *p, T v) { const T o = *p; *p = o | v; return o; }
template <class T> inline T atomicExch(T *p, T v) { const T o = *p; *p = v; return o; }
inline long long __double_as_longlong(double v) { long long b; memcpy(&b, &v, 8); return b; }
inline unsigned __float_as_uint(float v) { unsigned b; memcpy(&b, &v, 4); return b; }

// ---- integer intrinsics ----
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline unsigned __vcmpltu4(unsigned a, unsigned b) {         // per byte: 0xff where a < b (unsigned)
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) if (((a >> (8 * i)) & 0xffu) < ((b >> (8 * i)) & 0xffu)) r |= 0xffu << (8 * i);
    return r;
}
// per 16-bit half (DPX): max(a, b); min(a + b, c) with the sum wrapping in 16 bits, as the hardware does
inline unsigned __vmaxu2(unsigned a, unsigned b) {
    const unsigned lo = (a & 0xffffu) > (b & 0xffffu) ? (a & 0xffffu) : (b & 0xffffu), hi = (a >> 16) > (b >> 16) ? (a >> 16) : (b >> 16);
    return lo | (hi << 16);
}
inline unsigned __viaddmin_u16x2(unsigned a, unsigned b, unsigned c) {
    const unsigned slo = ((a & 0xffffu) + (b & 0xffffu)) & 0xffffu, shi = ((a >> 16) + (b >> 16)) & 0xffffu;
    const unsigned lo = slo < (c & 0xffffu) ? slo : (c & 0xffffu), hi = shi < (c >> 16) ? shi : (c >> 16);
    return lo | (hi << 16);
}
inline size_t __cvta_generic_to_global(const void *p) { return (size_t)p; }
// "shared window" addresses: 32-bit offsets from an anchor inside this library; the function-local statics that
// stand in for __shared__ arrays lie within 2 GB of it, so the offset survives the round trip through 32 bits
inline char *emu_shared_anchor() { static char anchor; return &anchor; }
inline size_t __cvta_generic_to_shared(const void *p) { return (size_t)((const char *)p - emu_shared_anchor()); }
inline unsigned long long *emu_shared_ptr(uint32_t addr) { return (unsigned long long *)(emu_shared_anchor() + (int32_t)addr); }

// ---- min / max as CUDA overloads them ----
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline unsigned min(int a, unsigned b) { return min((unsigned)a, b); }
inline unsigned min(unsigned a, int b) { return min(a, (unsigned)b); }
inline unsigned max(int a, unsigned b) { return max((unsigned)a, b); }
inline unsigned max(unsigned a, int b) { return max(a, (unsigned)b); }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }
inline long min(long a, long b) { return a < b ? a : b; }
inline long max(long a, long b) { return a > b ? a : b; }
inline float min(float a, float b) { return fminf(a, b); }
inline float max(float a, float b) { return fmaxf(a, b); }
