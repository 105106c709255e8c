// common.cuh -- shared declarations of libbonejthickness (sm_100a).  Internal; the public
// surface is include/bonej_thickness.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace bjt {

// EDT x-pass writes the in-row distance (not squared) as u16; rows without background carry GX_INF.
constexpr uint16_t GX_INF = 0xFFFFu;
// largest supported extent: 8 x 1024 voxels per row segment set of the x-pass kernel
constexpr int MAX_EXTENT = 8192;

struct StatsPartial {
    double sum, sum2;
    double mn, mx;
    long long count;
};

// sentinel of EDT_S1D: noResult = 3 (n+1)^2, n = max(w,h,d)      
__host__ __device__ inline uint32_t no_result(int n) { return 3u * (uint32_t)(n + 1) * (uint32_t)(n + 1); }

__device__ __forceinline__ uint32_t isqrt_floor(uint32_t v) {
    // exact floor(sqrt(v)) for v < 2^31
    uint32_t r = (uint32_t)sqrtf((float)v);
    while (r * r > v) --r;
    while ((r + 1u) * (r + 1u) <= v) ++r;
    return r;
}
// the same through the approximate square root (one MUFU): the float root of a value below 2^28 is off by far less
// than 1, so its truncation is the answer or one beside it
__device__ __forceinline__ uint32_t isqrt_fast(uint32_t v) {
#ifndef __CUDACC__   // a host compiler reading this header: the portable form
    uint32_t r = (uint32_t)sqrtf((float)v);
#else
    float f;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(f) : "f"((float)v));
    uint32_t r = (uint32_t)f;
#endif
    if (r * r > v) --r;
    else if ((r + 1u) * (r + 1u) <= v) ++r;
    return r;
}
// floor(sqrt(v)) for v < 2^18 without a correction step: sqrt(v + 0.5) lies at least 1 / (4 * 512) away from every
// integer, the approximate root (a couple of ulp of 2^-15) cannot cross one
__device__ __forceinline__ int isqrt_small(uint32_t v) {
#ifndef __CUDACC__   // a host compiler reading this header: the portable form
    return (int)sqrtf((float)v + 0.5f);
#else
    float f;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(f) : "f"((float)v + 0.5f));
    return (int)f;
#endif
}

// A shared-memory cell by its address in the shared window, and a max-reduction on it.  Through generic pointers every
// atomicMax of the disc painter recomputed the window base (S2UR / UMOV / ULEA + LEA) and sat in its own branch region;
// with the 32-bit address kept in a register it is one address add and one RED (callers that want the reduction off
// for a lane hand in the address of a scratch word instead: a select, where ptxas makes a predicated RED a branch).
#ifdef __CUDACC__
typedef uint32_t smem_addr_t;
__device__ __forceinline__ smem_addr_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void smem_red_max(smem_addr_t a, uint32_t v) {
    asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
#else   // a host compiler reading this header (the CPU emulation of the kernels): plain pointers
typedef uintptr_t smem_addr_t;
inline smem_addr_t smem_addr(const void *p) { return (uintptr_t)p; }
inline void smem_red_max(smem_addr_t a, uint32_t v) {
    uint32_t *q = (uint32_t *)a;
    if (*q < v) *q = v;
}
#endif

__device__ __forceinline__ uint32_t isqrt_ceil(uint32_t v) {
    uint32_t r = isqrt_floor(v);
    return r * r < v ? r + 1u : r;
}

// ---- launchers (each returns the number of kernels it launched; errors via cudaGetLastError) ----
int launch_edt_x(const uint8_t *in, int w, int h, int d, int inverse, int thresh, uint16_t *gx, cudaStream_t s);
int launch_edt_y(const uint16_t *gx, int w, int h, int d, uint32_t n0, uint32_t *gy, cudaStream_t s);
// present: n0+1 bytes, must be zeroed by the caller; maxval: one u32, zeroed by the caller
int launch_edt_z(const uint32_t *gy, int w, int h, int d, uint32_t n0, uint32_t *sq, uint8_t *present,
                 uint32_t *maxval, cudaStream_t s);

// distinct-value list from the presence table; values[] sized >= number of present values
int launch_collect_values(const uint8_t *present, uint32_t vmax, uint32_t *values, uint32_t *count, cudaStream_t s);
// t1/t2/t3 indexed by r^2 value (direct index)
int launch_ridge_template(const uint32_t *values, uint32_t nvalues, uint32_t *t1, uint32_t *t2, uint32_t *t3,
                          cudaStream_t s);
// same, compact output (parity test of createTemplate)
int launch_ridge_template_compact(const uint32_t *values, uint32_t nvalues, uint32_t *t1, uint32_t *t2,
                                  uint32_t *t3, cudaStream_t s);
int launch_mark_present(const uint32_t *sq, size_t n, uint8_t *present, uint32_t *maxval, cudaStream_t s);
int launch_ridge(const uint32_t *sq, int w, int h, int d, const uint32_t *t1, const uint32_t *t2,
                 const uint32_t *t3, uint32_t *ridge, unsigned long long *count, cudaStream_t s);

// brute-force scatter paint (path 2): points = compacted ridge (x | y<<16, z, r^2, pad)
int launch_compact_ridge(const uint32_t *ridge, int w, int h, int d, uint4 *points, unsigned long long *count,
                         cudaStream_t s);
int launch_paint_scatter(const uint4 *points, unsigned long long npoints, int w, int h, int d, uint32_t *rsq,
                         cudaStream_t s);
int launch_fill_u32(uint32_t *p, size_t n, uint32_t v, cudaStream_t s);

// fronts along z + discs in the plane (path 4), see paint_disc.cu
bool paint_disc_supports(uint32_t rsq_max);
void paint_disc_plan(int w, int h, int nplanes, int items_per_voxel, size_t budget_bytes, int *chunk_planes,
                     size_t *workspace_bytes);
int launch_paint_disc(const uint32_t *ridge, int w, int h, int dext, int z_lo, int z_hi, uint32_t rsq_max,
                      int chunk_planes, int items_per_voxel, void *workspace, uint32_t *out, uint32_t *flags,
                      cudaEvent_t *events, cudaStream_t s);

int launch_isqrt_small_check(uint32_t *mismatches, cudaStream_t s);   // self-check of the disc kernel's square root

// separable paint (paths 0/1), see paint_sep.cu
size_t paint_sep_workspace_bytes(int w, int h, int d, int wide);
void paint_sep_set_limits(int cap, int fmax);          // test hook: 0 = default
void paint_sep_set_xslab_cols(int cols);               // 0 = the library's choice (process-wide)
int paint_sep_natural_xslab_cols(int h, int d);        // the library's choice for slabs of h x d per column, unclipped
unsigned long long paint_sep_redo_lines(const void *workspace);
int launch_paint_sep(const uint32_t *ridge, int w, int h, int d, int wide, void *workspace, uint32_t *rsq,
                     uint32_t *overflow_flag, cudaStream_t s);
// the same painter in steps, for z-slab sharded runs (stages 1 and 2 are plane-local; only the z sweeps of
// stage 3 cross slab boundaries).  For every x-slab xs < paint_sep_nxslabs(w): stage2, export towards the
// neighbours, [exchange], stage3 with what the neighbours exported.
// Boundary staircases: per column (y * xw + lx) up to kcap items (x = reach beyond the receiver's first /
// last plane, y = tag), tags descending, and a count.
constexpr int SEP_MAX_SOURCES = 4;
struct SepImports {
    int n_lo, n_hi, kcap;                                   // sources below plane 0 / above plane d-1
    const uint32_t *lo_cnt[SEP_MAX_SOURCES]; const uint2 *lo_items[SEP_MAX_SOURCES];
    const uint32_t *hi_cnt[SEP_MAX_SOURCES]; const uint2 *hi_items[SEP_MAX_SOURCES];
};
int paint_sep_nxslabs(int w, int h, int d);
int paint_sep_xslab_width(int w, int h, int d, int xs);
int paint_sep_stage1(const uint32_t *ridge, int w, int h, int d, int variant, void *workspace, uint32_t *flags, cudaStream_t s);
int paint_sep_stage2(int w, int h, int d, int variant, void *workspace, int xs, uint32_t *flags, cudaStream_t s);
// side +1: items of the last `hh` planes that reach plane d - 1 + gap or beyond; side -1: items of the first
// `hh` planes that reach plane -gap or beyond (gap >= 1)
int paint_sep_export(int w, int h, int d, int variant, void *workspace, int xs, int side, int gap, int hh,
                     uint32_t *cnt, uint2 *items, int kcap, uint32_t *flags, cudaStream_t s);
int paint_sep_stage3(int w, int h, int d, int variant, void *workspace, int xs, const SepImports *imp, uint32_t *rsq,
                     uint32_t *flags, cudaStream_t s);

// 2*sqrt + cleanup + mask + NaN/calibrate + stats
// output / statistics restricted to planes [z_lo, z_hi) of the local volume; `out` starts at plane z_lo
// thick_tab: table of 2*sqrt(r^2) filled by launch_thickness_table (thickness_table_bytes() bytes), or null
int launch_finish(const uint32_t *rsq, const uint8_t *original, int w, int h, int d, int z_lo, int z_hi, int inverse,
                  int calibrate, float pixel_width, float *out, StatsPartial *partials, int npartials,
                  StatsPartial *result, const float *thick_tab, cudaStream_t s);
int launch_thickness_table(float *tab, cudaStream_t s);
size_t thickness_table_bytes();
int finish_partials(int w, int h, int nz);
int launch_stats(const float *map, size_t n, StatsPartial *partials, int npartials, StatsPartial *result,
                 cudaStream_t s);
// segmented statistics of a map on the device (segstats.cu); work = 5 * segments doubles: sum, max, count, mean, ssq
int launch_label_stats(const float *map, const int32_t *labels, size_t n, int nseg, const double *sizes, double *work,
                       uint32_t *bad, cudaStream_t s);
int launch_slice_stats(const float *map, int w, int h, int d, int rx, int ry, int rw, int rh, double *work, cudaStream_t s);
// FindRidgePoints (ridgepoints.cu)
int launch_pad_binary(const uint8_t *in, int w, int h, int d, uint8_t *out, cudaStream_t s);
int launch_ridge_points_map(const uint32_t *sq, int W, int H, int D, float *a, float *b, float *c, uint32_t *max_bits,
                            cudaStream_t s);
int launch_ridge_points_collect(const float *ridge, size_t n, float threshold, unsigned long long *out, unsigned long long cap,
                                unsigned long long *count, cudaStream_t s);
int launch_check_binary(const uint8_t *in, size_t n, uint32_t *bad, cudaStream_t s);

// strided copies by the SMs, all blocks of a batch in one launch (peer_copy.cu); units of 16 bytes
constexpr int COPY_BATCH_MAX = 16;
struct CopyDesc { const uint4 *src; uint4 *dst; size_t spitch16, dpitch16, width16, height; };
struct CopyBatch { int n; CopyDesc d[COPY_BATCH_MAX]; };
int launch_copy2d_batch(const CopyBatch &b, int sm_count, cudaStream_t s);

int launch_phantom(const int32_t *shapes, int nshapes, int w, int h, int d, int z0, int dz, uint8_t *out,
                   cudaStream_t s);

}  // namespace bjt
