// ridgepoints.cu -- seed points for Ellipsoid Factor: org.bonej.ops.skeletonize.FindRidgePoints
// (Modern/ops/src/main/java/org/bonej/ops/skeletonize/FindRidgePoints.java:55-116) on the CUDA EDT.
//
//   expanded = input with a one-voxel border of background            (Views.expandZero, :87-90)
//   dist     = Euclidean distance to the nearest background, float    (ops.image().distancetransform, :92)
//   open     = dilate(erode(dist)), close = erode(dilate(dist)), structuring element: the grid ball of radius 2
//              (HyperSphereShape(2): 33 offsets with dx^2+dy^2+dz^2 <= 4); neighbours outside the image do not
//              take part (the morphology ops extend with the type's max / min)                          (:94-98)
//   ridge    = close - open, set to 0 where open < 1 + 1e-12                                            (:99-113)
//   seeds    = voxels with ridge > (float)fraction * max(ridge), as (x - 0.5, y - 0.5, z - 0.5) in the
//              coordinates of the expanded image, i.e. voxel centres of the input                       (:60-82)
// The squared distances come from the library's own EDT (edt.cu); everything here is small stencils.
#include "common.cuh"
#include <math.h>

namespace bjt {

namespace {

// the 33 offsets of the radius-2 grid ball, handed to the kernels by value (kernel parameters are constant memory of
// the launch: nothing per device or per process to initialise)
struct Ball { int8_t o[33][3]; };

__global__ void pad_kernel(const uint8_t *__restrict__ in, int w, int h, int d, uint8_t *__restrict__ out) {
    const int W = w + 2, H = h + 2, D = d + 2;
    const size_t n = (size_t)W * H * D;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)((i / W) % H), z = (int)(i / ((size_t)W * H));
        uint8_t v = 0;
        if (x >= 1 && x <= w && y >= 1 && y <= h && z >= 1 && z <= d)
            v = in[((size_t)(z - 1) * h + (y - 1)) * w + (x - 1)] ? 255 : 0;
        out[i] = v;
    }
}

__global__ void sqrt_kernel(const uint32_t *__restrict__ sq, size_t n, float *__restrict__ dist) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dist[i] = (float)sqrt((double)sq[i]);
}

// grey erosion (IS_MAX = false) or dilation (IS_MAX = true) by the radius-2 ball, neighbours outside ignored
template <bool IS_MAX>
__global__ void morph_kernel(const float *__restrict__ src, int W, int H, int D, float *__restrict__ dst, const Ball ball) {
    const size_t n = (size_t)W * H * D;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)((i / W) % H), z = (int)(i / ((size_t)W * H));
        float v = IS_MAX ? -INFINITY : INFINITY;
#pragma unroll 1
        for (int k = 0; k < 33; ++k) {
            const int xx = x + ball.o[k][0], yy = y + ball.o[k][1], zz = z + ball.o[k][2];
            if (xx < 0 || yy < 0 || zz < 0 || xx >= W || yy >= H || zz >= D) continue;
            const float s = __ldg(src + ((size_t)zz * H + yy) * W + xx);
            v = IS_MAX ? fmaxf(v, s) : fminf(v, s);
        }
        dst[i] = v;
    }
}

// ridge = close - open (0 where open < 1 + 1e-12), running maximum (values are >= 0: float bits order like the numbers)
__global__ void ridge_value_kernel(const float *close, const float *__restrict__ open, size_t n, float *ridge,
                                   uint32_t *__restrict__ max_bits) {      // ridge may be the close buffer
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float r = close[i] - open[i];
        if ((double)open[i] < 1.0 + 1e-12) r = 0.f;
        ridge[i] = r;
        m = fmaxf(m, r);
    }
    for (int k = 16; k; k >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, k));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(max_bits, __float_as_uint(m));
}

__global__ void collect_kernel(const float *__restrict__ ridge, size_t n, float threshold, unsigned long long *__restrict__ out,
                               unsigned long long cap, unsigned long long *__restrict__ count) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if ((double)ridge[i] <= (double)threshold) continue;
        const unsigned long long slot = atomicAdd(count, 1ull);
        if (slot < cap) out[slot] = (unsigned long long)i;
    }
}

unsigned grid_for(size_t n) {
    const size_t b = (n + 255) / 256;
    return (unsigned)(b < 148 * 32 ? (b ? b : 1) : 148 * 32);
}

Ball make_ball() {
    Ball b;
    int k = 0;
    for (int z = -2; z <= 2; ++z)
        for (int y = -2; y <= 2; ++y)
            for (int x = -2; x <= 2; ++x)
                if (x * x + y * y + z * z <= 4) { b.o[k][0] = (int8_t)x; b.o[k][1] = (int8_t)y; b.o[k][2] = (int8_t)z; ++k; }
    return b;
}

}  // namespace

int launch_pad_binary(const uint8_t *in, int w, int h, int d, uint8_t *out, cudaStream_t s) {
    pad_kernel<<<grid_for((size_t)(w + 2) * (h + 2) * (d + 2)), 256, 0, s>>>(in, w, h, d, out);
    return 1;
}

// sq: squared EDT of the expanded image (W x H x D); a, b, c: three float scratch volumes; on return `a` holds the ridge.
int launch_ridge_points_map(const uint32_t *sq, int W, int H, int D, float *a, float *b, float *c, uint32_t *max_bits,
                            cudaStream_t s) {
    const Ball ball = make_ball();
    const size_t n = (size_t)W * H * D;
    const unsigned g = grid_for(n);
    sqrt_kernel<<<g, 256, 0, s>>>(sq, n, a);                       // a = dist
    morph_kernel<false><<<g, 256, 0, s>>>(a, W, H, D, b, ball);          // b = erode(dist)
    morph_kernel<true><<<g, 256, 0, s>>>(b, W, H, D, c, ball);           // c = open
    morph_kernel<true><<<g, 256, 0, s>>>(a, W, H, D, b, ball);           // b = dilate(dist)
    morph_kernel<false><<<g, 256, 0, s>>>(b, W, H, D, a, ball);          // a = close
    cudaMemsetAsync(max_bits, 0, 4, s);
    ridge_value_kernel<<<g, 256, 0, s>>>(a, c, n, a, max_bits);    // a = ridge
    return 6;
}

int launch_ridge_points_collect(const float *ridge, size_t n, float threshold, unsigned long long *out, unsigned long long cap,
                                unsigned long long *count, cudaStream_t s) {
    cudaMemsetAsync(count, 0, 8, s);
    collect_kernel<<<grid_for(n), 256, 0, s>>>(ridge, n, threshold, out, cap, count);
    return 1;
}

}  // namespace bjt
