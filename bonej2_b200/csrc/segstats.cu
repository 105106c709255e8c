// segstats.cu -- segmented statistics of a thickness map that is already on the device: the reductions the
// Legacy consumers of the map run on the CPU after Thickness.
//
//   per label   ParticleAnalysis.getMeanStdDev (Legacy/bonej/.../ParticleAnalysis.java:325-364): over the voxels
//               of particle p with value > 0:  mean = sum / size_p (the particle's voxel count, zero-valued and NaN
//               voxels included in the divisor), sd = sqrt(sum (v - mean)^2 / size_p), max.  Label 0 is background
//               and stays 0.
//   per slice   SliceGeometry (Legacy/bonej/.../SliceGeometry.java:876-900, 1172-1200): over the pixels of slice z
//               inside the ROI rectangle with value > 0: mean = sum / count, max, sd = sqrt(sum (mean - v)^2 / count).
//
// Both are two passes in the reference (the residuals use the finished mean); so are they here: pass 1
// accumulates (count, sum, max) per segment, pass 2 the squared residuals, all in double.  A warp first
// combines the lanes that hit the same segment (__match_any_sync), so a run of equal labels costs one atomic.
#include "common.cuh"

namespace bjt {

namespace {

__device__ __forceinline__ void atomic_max_double(double *addr, double v) {
    // values are >= 0: the bit patterns order like the numbers
    atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}

// segment-wise warp combine of (x, m) for lanes with the same key; the lowest lane of each group adds
template <bool WITH_MAX>
__device__ __forceinline__ void seg_add(unsigned active, int key, double x, double m, double cnt, double *sum, double *mx,
                                        double *count) {
    const unsigned peers = __match_any_sync(active, key);
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    double sx = 0.0, sm = 0.0, sc = 0.0;
    for (unsigned rest = peers; rest;) {                       // sum in lane order: the same grouping gives the same result
        const int l = __ffs(rest) - 1;
        rest &= rest - 1;
        sx += __shfl_sync(peers, x, l);
        sc += __shfl_sync(peers, cnt, l);
        if (WITH_MAX) sm = fmax(sm, __shfl_sync(peers, m, l));
    }
    if (lane == leader) {
        atomicAdd(sum + key, sx);
        if (count) atomicAdd(count + key, sc);
        if (WITH_MAX) atomic_max_double(mx + key, sm);
    }
}

// pass 1: acc[0][k] = sum, acc[1][k] = max, acc[2][k] = count (values > 0 only)
__global__ void __launch_bounds__(256) label_pass1(const float *__restrict__ map, const int32_t *__restrict__ labels, size_t n,
                                                    int nseg, double *sum, double *mx, double *count, uint32_t *bad) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = map[i];
        const int k = labels[i];
        const bool pos = v > 0.0f, known = k >= 0 && k < nseg;
        if (pos && !known) atomicExch(bad, 1u);
        const bool in = pos && known;
        const unsigned active = __ballot_sync(__activemask(), in);
        if (in) seg_add<true>(active, k, (double)v, (double)v, 1.0, sum, mx, count);
    }
}

__global__ void __launch_bounds__(256) label_pass2(const float *__restrict__ map, const int32_t *__restrict__ labels, size_t n,
                                                    int nseg, const double *__restrict__ mean, double *ssq) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = map[i];
        const int k = labels[i];
        const bool in = v > 0.0f && k >= 0 && k < nseg;
        const unsigned active = __ballot_sync(__activemask(), in);
        if (in) {
            const double r = (double)v - mean[k];
            seg_add<false>(active, k, r * r, 0.0, 0.0, ssq, nullptr, nullptr);
        }
    }
}

// one block per (slice, row group): rows of the ROI of one slice, block-reduced, one atomic per block
template <int PASS>
__global__ void __launch_bounds__(256) slice_pass(const float *__restrict__ map, int w, int h, int rx, int ry, int rw, int rh,
                                                   int rows_per_block, const double *__restrict__ mean, double *sum, double *mx,
                                                   double *count, double *ssq) {
    const int z = blockIdx.y;
    const int y0 = ry + blockIdx.x * rows_per_block, y1 = min(ry + rh, y0 + rows_per_block);
    const float *plane = map + (size_t)z * w * h;
    double s = 0.0, m = 0.0, c = 0.0;
    const double mu = PASS == 2 ? mean[z] : 0.0;
    for (int y = y0; y < y1; ++y)
        for (int x = rx + threadIdx.x; x < rx + rw; x += blockDim.x) {
            const float v = plane[(size_t)y * w + x];
            if (v > 0.0f) {
                if (PASS == 1) { s += (double)v; m = fmax(m, (double)v); c += 1.0; }
                else { const double r = mu - (double)v; s += r * r; }
            }
        }
    __shared__ double sh[3][8];
    for (int k = 16; k; k >>= 1) {
        s += __shfl_down_sync(0xFFFFFFFFu, s, k);
        c += __shfl_down_sync(0xFFFFFFFFu, c, k);
        m = fmax(m, __shfl_down_sync(0xFFFFFFFFu, m, k));
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sh[0][warp] = s; sh[1][warp] = m; sh[2][warp] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) { s += sh[0][k]; m = fmax(m, sh[1][k]); c += sh[2][k]; }
        if (PASS == 1) {
            if (c > 0.0) { atomicAdd(sum + z, s); atomicAdd(count + z, c); atomic_max_double(mx + z, m); }
        } else if (s != 0.0) {
            atomicAdd(ssq + z, s);
        }
    }
}

// mean[k] = sum[k] / divisor[k]  (divisor: particle sizes, or the pass-1 counts)
__global__ void seg_mean(const double *sum, const double *divisor, int nseg, double *mean) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nseg) mean[k] = sum[k] / divisor[k];
}

}  // namespace

// work: 5 * nseg doubles (sum, max, count, mean, ssq), zeroed here.  sizes: nseg doubles (particle voxel counts).
int launch_label_stats(const float *map, const int32_t *labels, size_t n, int nseg, const double *sizes, double *work,
                       uint32_t *bad, cudaStream_t s) {
    double *sum = work, *mx = work + nseg, *count = work + 2 * (size_t)nseg, *mean = work + 3 * (size_t)nseg,
           *ssq = work + 4 * (size_t)nseg;
    cudaMemsetAsync(work, 0, 5 * (size_t)nseg * sizeof(double), s);
    const unsigned grid = (unsigned)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    label_pass1<<<grid ? grid : 1, 256, 0, s>>>(map, labels, n, nseg, sum, mx, count, bad);
    seg_mean<<<(nseg + 255) / 256, 256, 0, s>>>(sum, sizes, nseg, mean);
    label_pass2<<<grid ? grid : 1, 256, 0, s>>>(map, labels, n, nseg, mean, ssq);
    return 3;
}

// work: 5 * d doubles (sum, max, count, mean, ssq), zeroed here
int launch_slice_stats(const float *map, int w, int h, int d, int rx, int ry, int rw, int rh, double *work, cudaStream_t s) {
    double *sum = work, *mx = work + d, *count = work + 2 * (size_t)d, *mean = work + 3 * (size_t)d, *ssq = work + 4 * (size_t)d;
    cudaMemsetAsync(work, 0, 5 * (size_t)d * sizeof(double), s);
    const int rows_per_block = 32;
    const dim3 grid((rh + rows_per_block - 1) / rows_per_block, d);
    slice_pass<1><<<grid, 256, 0, s>>>(map, w, h, rx, ry, rw, rh, rows_per_block, nullptr, sum, mx, count, nullptr);
    seg_mean<<<(d + 255) / 256, 256, 0, s>>>(sum, count, d, mean);
    slice_pass<2><<<grid, 256, 0, s>>>(map, w, h, rx, ry, rw, rh, rows_per_block, mean, nullptr, nullptr, nullptr, ssq);
    return 3;
}

}  // namespace bjt
