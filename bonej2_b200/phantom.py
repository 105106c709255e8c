"""Synthetic binary trabecular phantom (random rods + plates), the workload of BASELINE.json's
configs.  This module only generates the integer shape records; the
rasteriser is the CUDA kernel `bjt_phantom_u8` (bonej2_b200/csrc/phantom.cu).  The records use
quarter-voxel fixed point and every inside test is exact int64 arithmetic, so any rasteriser
that follows the record definition produces the same bytes.

Recipe per 256^3 of volume: 200 rods (capsule radius U[3,8] voxels, length U[0.3,0.9]*256,
uniform start point and orientation) + 40 plates (disc radius U[40,90], half-thickness U[2,5],
uniform centre and normal).  Feature sizes stay fixed in voxels as the volume grows.
RNG = splitmix64 so other languages can reproduce the records.
"""
from __future__ import annotations

import math

import numpy as np

Q = 4          # fixed-point scale (quarter voxels)
REC = 16       # int32 per shape record
_M64 = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & _M64

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & _M64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
        return z ^ (z >> 31)

    def uniform(self, lo: float = 0.0, hi: float = 1.0) -> float:
        return lo + (hi - lo) * ((self.next() >> 11) * (1.0 / (1 << 53)))

    def direction(self):
        """uniform direction by rejection from the cube"""
        while True:
            x, y, z = self.uniform(-1, 1), self.uniform(-1, 1), self.uniform(-1, 1)
            n2 = x * x + y * y + z * z
            if 1e-3 < n2 <= 1.0:
                n = math.sqrt(n2)
                return x / n, y / n, z / n


def _clip(v, hi):
    return max(0, min(hi - 1, v))


def phantom_shapes(w: int, h: int, d: int, seed: int = 1, rods_per_256: float = 200.0,
                   plates_per_256: float = 40.0) -> np.ndarray:
    """Integer shape records [n, 16] (see oracle/phantom_cpu.c / csrc/phantom.cu for the layout)."""
    rng = SplitMix64(seed)
    scale = (w * h * d) / float(256 ** 3)
    n_rods = max(1, int(round(rods_per_256 * scale)))
    n_plates = max(1, int(round(plates_per_256 * scale)))
    rec = np.zeros((n_rods + n_plates, REC), np.int32)
    k = 0
    for _ in range(n_rods):
        ax, ay, az = rng.uniform(0, w), rng.uniform(0, h), rng.uniform(0, d)
        ux, uy, uz = rng.direction()
        length = rng.uniform(0.3, 0.9) * 256.0
        r = rng.uniform(3.0, 8.0)
        bx, by, bz = ax + ux * length, ay + uy * length, az + uz * length
        a = [int(round(c * Q)) for c in (ax, ay, az)]
        b = [int(round(c * Q)) for c in (bx, by, bz)]
        rq = int(round(r * Q))
        pad = (rq + Q - 1) // Q + 1
        lo = [min(a[i], b[i]) // Q - pad for i in range(3)]
        hi = [max(a[i], b[i]) // Q + pad + 1 for i in range(3)]
        rec[k, 0] = 0
        rec[k, 1:4] = a
        rec[k, 4:7] = b
        rec[k, 7] = rq
        rec[k, 8] = 0
        rec[k, 9:15] = [_clip(lo[0], w), _clip(hi[0], w), _clip(lo[1], h), _clip(hi[1], h),
                        _clip(lo[2], d), _clip(hi[2], d)]
        k += 1
    for _ in range(n_plates):
        cx, cy, cz = rng.uniform(0, w), rng.uniform(0, h), rng.uniform(0, d)
        nx, ny, nz = rng.direction()
        rad = rng.uniform(40.0, 90.0)
        t = rng.uniform(2.0, 5.0)
        c = [int(round(v * Q)) for v in (cx, cy, cz)]
        m = [int(round(v * 1024)) for v in (nx, ny, nz)]
        rq, tq = int(round(rad * Q)), int(round(t * Q))
        pad = (rq + Q - 1) // Q + 2
        lo = [c[i] // Q - pad for i in range(3)]
        hi = [c[i] // Q + pad + 1 for i in range(3)]
        rec[k, 0] = 1
        rec[k, 1:4] = c
        rec[k, 4:7] = m
        rec[k, 7] = rq
        rec[k, 8] = tq
        rec[k, 9:15] = [_clip(lo[0], w), _clip(hi[0], w), _clip(lo[1], h), _clip(hi[1], h),
                        _clip(lo[2], d), _clip(hi[2], d)]
        k += 1
    return rec
