"""What pruning ridge points whose ball lies inside the ball of another ridge point within k voxels would
leave (exact: a contained ball can never be the largest one anywhere).  Writes the pruned ridge volume.
Continuous-ball test (sufficient for the voxel balls): R' - R - delta^2 >= 0 and 4 R delta^2 <= (R' - R - delta^2)^2.
usage: python tools/study/prune_ridge.py in.u32 n k out.u32"""
import sys
import numpy as np
src, n, k, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
r = np.fromfile(src, np.uint32).reshape(n, n, n).astype(np.int64)
keep = r > 0
P = np.pad(r, k)
for dz in range(-k, k + 1):
    for dy in range(-k, k + 1):
        for dx in range(-k, k + 1):
            d2 = dx * dx + dy * dy + dz * dz
            if d2 == 0:
                continue
            nb = P[k + dz:k + dz + n, k + dy:k + dy + n, k + dx:k + dx + n]
            D = nb - r - d2
            keep &= ~((D >= 0) & (4 * r * d2 <= D * D))
print(f"k={k}: ridge points {int((r > 0).sum())} -> {int(keep.sum())}")
np.where(keep, r, 0).astype(np.uint32).tofile(out)
