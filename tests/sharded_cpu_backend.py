"""Oracle-backed stage backend for bonej2_b200.sharded -- TEST ONLY.  It lets the z-slab exchange logic
(transposes, halo gathers, statistics combination) run on CPU tensors over gloo, with the CPU oracle
doing the arithmetic of each stage on the pieces a rank holds."""
import numpy as np
import torch

from bonej2_b200.sharded import PartialStats
from oracle import lt_oracle


def _np(t):
    return t.numpy()


class OracleStages:
    device = torch.device("cpu")

    def edt_x(self, slab_u8, inverse, threshold=128):
        v = _np(slab_u8)
        bg = (v < threshold) ^ bool(inverse)
        dz, h, w = v.shape
        x = np.arange(w)
        dist = np.abs(x[None, :] - x[:, None]).astype(np.int64)          # [x][x']
        big = 1 << 40
        g = np.where(bg[:, :, None, :], dist[None, None, :, :], big).min(axis=3)
        out = np.where(g >= big, 0xFFFF, g).astype(np.uint16)
        return torch.from_numpy(out.view(np.int16).copy())

    def edt_y(self, gx, n_sentinel):
        n0 = 3 * (n_sentinel + 1) ** 2
        g = _np(gx).view(np.uint16).astype(np.int64)
        f = np.where(g == 0xFFFF, n0, g * g)                              # [z][y'][x]
        h = f.shape[1]
        y = np.arange(h)
        d2 = ((y[:, None] - y[None, :]) ** 2).astype(np.int64)           # [y][y']
        out = np.minimum((f[:, None, :, :] + d2[None, :, :, None]).min(axis=2), n0)
        return torch.from_numpy(out.astype(np.int32))

    def edt_z(self, gy_t, n_sentinel):
        n0 = 3 * (n_sentinel + 1) ** 2
        f = _np(gy_t).astype(np.int64)                                     # [z'][y][x]
        d = f.shape[0]
        if f.size == 0:
            return torch.from_numpy(f.astype(np.int32))
        z = np.arange(d)
        d2 = ((z[:, None] - z[None, :]) ** 2).astype(np.int64)
        out = np.minimum((f[None, :, :, :] + d2[:, :, None, None]).min(axis=1), n0)
        return torch.from_numpy(out.astype(np.int32))

    def max_value(self, t):
        return int(t.max().item()) if t.numel() else 0

    def ridge(self, sq_ext):
        sq = _np(sq_ext).astype(np.uint32)
        dist = np.sqrt(sq.astype(np.float64)).astype(np.float32)
        rg = lt_oracle.ridge(dist)
        return torch.from_numpy(np.where(rg > 0, sq, 0).astype(np.int32))

    def paint(self, ridge_ext):
        rg = _np(ridge_ext).astype(np.uint32)
        _, rsq = lt_oracle.paint(np.sqrt(rg.astype(np.float64)).astype(np.float32), nthreads=2)
        return torch.from_numpy(rsq.astype(np.int32))

    def paint_range(self, ridge_ext, z_lo, z_hi, max_rsq):
        return self.paint(ridge_ext)[z_lo:z_hi]

    def full(self, shape, value):
        return torch.full(shape, value, dtype=torch.int32)

    def finish(self, rsq_sub, orig_core_u8, z_lo, z_hi, inverse, mask, pixel_width):
        R = _np(rsq_sub).astype(np.float64)
        lt = (2.0 * np.sqrt(R)).astype(np.float32)
        cl = lt_oracle.cleanup(lt)[z_lo:z_hi]
        if mask:
            cl = lt_oracle.mask(_np(orig_core_u8), cl, inverse=inverse)
        m = lt_oracle.calibrate(cl, pixel_width)
        ok = ~np.isnan(m)
        v = m[ok].astype(np.float64)
        part = PartialStats(int(ok.sum()), float(v.sum()), float((v * v).sum()),
                            float(v.min()) if v.size else float("inf"), float(v.max()) if v.size else float("-inf"))
        return torch.from_numpy(m), part
