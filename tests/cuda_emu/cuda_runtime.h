/*
 * cuda_runtime.h -- TEST ONLY: a small CPU emulation of the CUDA constructs this repository's kernels use, so the
 * kernels' own source (bonej2_b200/csrc/*.cu, rewritten only at their <<< >>> launches by tests/cuda_emu/build.py)
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
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
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
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->ms - a->ms); return cudaSuccess; }
template <class T> inline cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)aligned_alloc(256, (n + 255) / 256 * 256); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
template <class T> inline cudaError_t cudaMallocHost(T **p, size_t n) { *p = (T *)aligned_alloc(256, (n + 255) / 256 * 256); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemGetInfo(size_t *free_b, size_t *total_b) { *free_b = (size_t)64 << 30; *total_b = (size_t)64 << 30; return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(dst, src, n); return cudaSuccess; }

#define cudaFuncSetAttribute(...) 0
#define cudaMemsetAsync(p, v, n, s) (memset((p), (v), (n)), 0)
#define cudaMemcpyToSymbol(sym, src, n) (memcpy(&(sym), (src), (n)), 0)
#define cudaMemcpy(dst, src, n, kind) (memcpy((dst), (src), (n)), 0)

namespace emu {

constexpr size_t STACK_BYTES = 512 * 1024;

struct Warp {
    unsigned long long vals[32], snap[32];
    bool live[32];
    int alive = 0, arrived = 0;
    unsigned gen = 0;
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

inline void warp_release_if_complete(Warp &w) {
    if (w.alive > 0 && w.arrived == w.alive) {
        memcpy(w.snap, w.vals, sizeof w.snap);
        w.arrived = 0;
        ++w.gen;
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
    w.live[f.lin & 31] = false;
    --w.alive;
    w.vals[f.lin & 31] = 0;
    warp_release_if_complete(w);
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

// every live lane of the warp contributes one 64-bit value; returns the warp's values (dead lanes: 0)
inline const unsigned long long *warp_gather(unsigned long long v) {
    Block &b = *blk();
    const int lin = cur().lin;
    Warp &w = b.warps[lin >> 5];
    const unsigned g = w.gen;
    w.vals[lin & 31] = v;
    ++w.arrived;
    warp_release_if_complete(w);
    while (w.gen == g) yield();
    return w.snap;
}

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
        for (auto &w : b.warps) { w.alive = 32; w.arrived = 0; for (int l = 0; l < 32; ++l) { w.live[l] = true; w.vals[l] = 0; } }
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

template <class T> inline T __shfl_sync(unsigned, T v, int src) {
    const unsigned long long *s = emu::warp_gather(emu_bits(v));
    return emu_unbits<T>(s[src & 31]);
}
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned delta) {
    const int lane = emu_lane();
    const unsigned long long *s = emu::warp_gather(emu_bits(v));
    return lane >= (int)delta ? emu_unbits<T>(s[lane - (int)delta]) : v;
}
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned delta) {
    const int lane = emu_lane();
    const unsigned long long *s = emu::warp_gather(emu_bits(v));
    return lane + (int)delta < 32 ? emu_unbits<T>(s[lane + (int)delta]) : v;
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
    const int lane = emu_lane();
    const unsigned long long *s = emu::warp_gather(emu_bits(v));
    return emu_unbits<T>(s[(lane ^ lane_mask) & 31]);
}
inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
inline unsigned __ballot_sync(unsigned, int pred) {
    const unsigned long long *s = emu::warp_gather(pred ? 1ull : 0ull);
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) m |= s[l] ? (1u << l) : 0u;
    return m;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0u; }
inline unsigned __reduce_max_sync(unsigned, unsigned v) {
    const unsigned long long *s = emu::warp_gather(v);
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) m = (unsigned)s[l] > m ? (unsigned)s[l] : m;
    return m;
}
inline unsigned __reduce_add_sync(unsigned, unsigned v) {
    const unsigned long long *s = emu::warp_gather(v);
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) m += (unsigned)s[l];
    return m;
}

// ---- atomics: one OS thread, so plain read-modify-write ----
template <class T> inline T atomicAdd(T *p, T v) { const T o = *p; *p = o + v; return o; }
template <class T> inline T atomicMax(T *p, T v) { const T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicOr(T *p, T v) { const T o = *p; *p = o | v; return o; }

// ---- integer intrinsics ----
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline unsigned __vcmpltu4(unsigned a, unsigned b) {         // per byte: 0xff where a < b (unsigned)
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) if (((a >> (8 * i)) & 0xffu) < ((b >> (8 * i)) & 0xffu)) r |= 0xffu << (8 * i);
    return r;
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
