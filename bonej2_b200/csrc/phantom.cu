// phantom.cu -- rasteriser of the synthetic trabecular phantom (random rods + plates), sm_100a.
// Workload definition: SURVEY.md 8(d); record layout: bonej2_b200/phantom.py.  One block per shape
// walks the shape's bounding box; inside tests are exact int64 arithmetic in quarter-voxel fixed
// point, so the bytes equal those of any other rasteriser following the record definition.
#include "common.cuh"

namespace bjt {

constexpr int PH_Q = 4, PH_REC = 16;

__device__ __forceinline__ bool inside_shape(const int32_t *s, long long px, long long py, long long pz) {
    const long long ux = px - s[1], uy = py - s[2], uz = pz - s[3];
    const long long u2 = ux * ux + uy * uy + uz * uz;
    const long long r2 = (long long)s[7] * s[7];
    if (s[0] == 0) {   // rod: capsule a-b
        const long long ex = (long long)s[4] - s[1], ey = (long long)s[5] - s[2], ez = (long long)s[6] - s[3];
        const long long L2 = ex * ex + ey * ey + ez * ez;
        const long long dot = ux * ex + uy * ey + uz * ez;
        if (dot <= 0) return u2 <= r2;
        if (dot >= L2) {
            const long long vx = px - s[4], vy = py - s[5], vz = pz - s[6];
            return vx * vx + vy * vy + vz * vz <= r2;
        }
        return u2 * L2 - dot * dot <= r2 * L2;
    }
    // plate: disc, integer normal m
    const long long mx = s[4], my = s[5], mz = s[6];
    const long long M2 = mx * mx + my * my + mz * mz;
    const long long sd = ux * mx + uy * my + uz * mz;
    const long long t2 = (long long)s[8] * s[8];
    if (sd * sd > t2 * M2) return false;
    return u2 * M2 - sd * sd <= r2 * M2;
}

__global__ void __launch_bounds__(256) phantom_kernel(const int32_t *__restrict__ shapes, int nshapes, int w, int h,
                                                      int z0, int dz, uint8_t *__restrict__ out) {
    __shared__ int32_t s[PH_REC];
    const int si = blockIdx.x;
    if (si >= nshapes) return;
    if (threadIdx.x < PH_REC) s[threadIdx.x] = shapes[(size_t)si * PH_REC + threadIdx.x];
    __syncthreads();
    const int x0 = s[9], x1 = s[10], y0 = s[11], y1 = s[12];
    const int za = max(s[13], z0), zb = min(s[14], z0 + dz - 1);
    if (za > zb) return;
    const long long nx = x1 - x0 + 1, ny = y1 - y0 + 1, nz = zb - za + 1;
    const long long total = nx * ny * nz;
    for (long long i = (long long)blockIdx.y * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.y * blockDim.x) {
        const int x = x0 + (int)(i % nx);
        const int y = y0 + (int)((i / nx) % ny);
        const int z = za + (int)(i / (nx * ny));
        if (inside_shape(s, (long long)x * PH_Q, (long long)y * PH_Q, (long long)z * PH_Q))
            out[((size_t)(z - z0) * h + y) * w + x] = 255;
    }
}

int launch_phantom(const int32_t *shapes, int nshapes, int w, int h, int d, int z0, int dz, uint8_t *out,
                   cudaStream_t s) {
    (void)d;
    cudaMemsetAsync(out, 0, (size_t)w * h * dz, s);
    if (nshapes <= 0) return 0;
    dim3 grid((unsigned)nshapes, 8);
    phantom_kernel<<<grid, 256, 0, s>>>(shapes, nshapes, w, h, z0, dz, out);
    return 1;
}

}  // namespace bjt
