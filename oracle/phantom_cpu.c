/*
 * phantom_cpu.c -- CPU rasteriser for the synthetic trabecular phantom.  TEST INFRASTRUCTURE
 * (lives beside the oracle; the product's rasteriser is the CUDA kernel in
 * bonej2_b200/csrc/phantom.cu and must produce the same bytes -- tests/test_gpu_phantom.py).
 *
 * Shapes arrive as integer records generated on the host (bonej2_b200/phantom.py, splitmix64),
 * 16 x int32 each, coordinates in quarter-voxel fixed point (Q = 4):
 *   [0] type 0 = rod (capsule a-b, radius r), 1 = plate (disc centre a, integer normal m = b,
 *       radius r, half-thickness t)
 *   [1..3] a   [4..6] b   [7] r   [8] t   [9..14] bbox x0,x1,y0,y1,z0,z1 (voxels, inclusive)
 * All inside tests are exact int64 arithmetic so CPU and GPU agree bit for bit.
 * Workload definition: SURVEY.md section 8(d).
 */
#include <stdint.h>
#include <string.h>

#define PH_Q 4
#define PH_REC 16

static inline int inside_shape(const int32_t *s, int64_t px, int64_t py, int64_t pz) {
    const int64_t ux = px - s[1], uy = py - s[2], uz = pz - s[3];
    const int64_t u2 = ux * ux + uy * uy + uz * uz;
    const int64_t r2 = (int64_t)s[7] * s[7];
    if (s[0] == 0) {
        const int64_t ex = (int64_t)s[4] - s[1], ey = (int64_t)s[5] - s[2], ez = (int64_t)s[6] - s[3];
        const int64_t L2 = ex * ex + ey * ey + ez * ez;
        const int64_t dot = ux * ex + uy * ey + uz * ez;
        if (dot <= 0) return u2 <= r2;
        if (dot >= L2) {
            const int64_t vx = px - s[4], vy = py - s[5], vz = pz - s[6];
            return vx * vx + vy * vy + vz * vz <= r2;
        }
        return u2 * L2 - dot * dot <= r2 * L2;
    } else {
        const int64_t mx = s[4], my = s[5], mz = s[6];
        const int64_t M2 = mx * mx + my * my + mz * mz;
        const int64_t sd = ux * mx + uy * my + uz * mz;
        const int64_t t2 = (int64_t)s[8] * s[8];
        if (sd * sd > t2 * M2) return 0;
        return u2 * M2 - sd * sd <= r2 * M2;
    }
}

int lto_phantom_raster(const int32_t *shapes, int nshapes, int w, int h, int d, uint8_t *out) {
    memset(out, 0, (size_t)w * h * d);
    for (int i = 0; i < nshapes; i++) {
        const int32_t *s = shapes + (size_t)i * PH_REC;
        for (int z = s[13]; z <= s[14]; z++)
            for (int y = s[11]; y <= s[12]; y++) {
                uint8_t *row = out + ((size_t)z * h + y) * w;
                for (int x = s[9]; x <= s[10]; x++)
                    if (!row[x] && inside_shape(s, (int64_t)x * PH_Q, (int64_t)y * PH_Q, (int64_t)z * PH_Q)) row[x] = 255;
            }
    }
    return 0;
}
