"""How many disc rows a containment test against the items a few lanes away would remove (CPU study for the disc
painter, paint_disc.cu): per plane the Pareto fronts of (slack, tag) per voxel are formed as zfront_kernel does,
laid out in bucket order (row, 32-column bucket, voxel, tag descending) and every item is tested against the items
up to K positions before and behind it in its bucket: dropped when one of them, at another column of the same row,
has at least its tag and rho_j >= rho_i + 2 |dx| floor(sqrt(rho_i)) + dx^2.
usage: python tools/study/lane_prune_study.py ridge.u32 n [planes]"""
import sys
import numpy as np
src, n = sys.argv[1], int(sys.argv[2])
planes = [int(p) for p in sys.argv[3].split(",")] if len(sys.argv) > 3 else [n // 4, n // 2, 3 * n // 4]
r = np.fromfile(src, np.uint32).reshape(n, n, n).astype(np.int64)
emax = int(np.sqrt(r.max()))
tot_rows = 0
drop_rows = {k: 0 for k in (1, 2, 3, 4, 6, 8, 31)}
tot_items = 0
for z in planes:
    cands = []
    for dz in range(-emax, emax + 1):
        zz = z + dz
        if 0 <= zz < n:
            R = r[zz]
            s = np.where((R > 0) & (R >= dz * dz), R - dz * dz, -1)
            cands.append(np.stack([R, s]))
    C = np.stack(cands)                                   # [cand][2][y][x]
    R, S = C[:, 0], C[:, 1]
    R = np.where(S >= 0, R, 0)
    order = np.argsort(-R, axis=0, kind="stable")
    R = np.take_along_axis(R, order, 0); S = np.take_along_axis(S, order, 0)
    items = []                                            # (y, x, R, rho)
    best = np.full(R.shape[1:], -1)
    # walk in descending tag order; an item when the slack improves; same tag: the larger slack replaces
    keepR = []; keepS = []
    for c in range(R.shape[0]):
        imp = (S[c] > best) & (R[c] > 0)
        best = np.where(imp, S[c], best)
        keepR.append(np.where(imp, R[c], 0)); keepS.append(np.where(imp, S[c], -1))
    KR, KS = np.stack(keepR), np.stack(keepS)
    # same tag, more slack replaces the previous: drop an item followed by an item of the same tag
    for c in range(KR.shape[0] - 1, 0, -1):
        pass
    ys, xs = np.nonzero((KR > 0).any(0))
    for y, x in zip(ys, xs):
        col = [(int(KR[c, y, x]), int(KS[c, y, x])) for c in range(KR.shape[0]) if KR[c, y, x] > 0]
        front = {}
        for Rv, Sv in col:
            front[Rv] = max(front.get(Rv, -1), Sv)
        for Rv in sorted(front, reverse=True):
            items.append((y, x, Rv, front[Rv]))
    items.sort(key=lambda t: (t[0], t[1] // 32, t[1], -t[2]))
    A = np.array(items, np.int64)
    tot_items += len(A)
    rows = 2 * np.floor(np.sqrt(A[:, 3])).astype(np.int64) + 1
    tot_rows += int(rows.sum())
    bucket = A[:, 0] * 4096 + A[:, 1] // 32
    a2 = 2 * np.floor(np.sqrt(A[:, 3])).astype(np.int64)
    for K in drop_rows:
        dropped = np.zeros(len(A), bool)
        for k in range(1, K + 1):
            for sgn in (1, -1):
                j = np.arange(len(A)) + sgn * k
                ok = (j >= 0) & (j < len(A))
                jj = np.clip(j, 0, len(A) - 1)
                same = ok & (bucket[jj] == bucket)
                dx = np.abs(A[jj, 1] - A[:, 1])
                dropped |= same & (dx != 0) & (A[jj, 2] >= A[:, 2]) & (A[jj, 3] >= A[:, 3] + a2 * dx + dx * dx)
        drop_rows[K] += int(rows[dropped].sum())
print("items", tot_items, "disc rows", tot_rows)
for K, v in drop_rows.items():
    print(f"window +-{K}: rows dropped {100.0 * v / tot_rows:.1f} %")
