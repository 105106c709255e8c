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

    # the painter in steps, on the CPU model of the separable painter (same calls and buffer layout as the
    # CUDA library's bjt_dev_paint_open .. _close), column blocks included: like the library, a rank cuts the
    # width into blocks of its own choosing unless a width has been set -- block_rule is the stand-in for the
    # library's "blocks sized by the slab's volume" (thicker slab, narrower blocks)
    block_rule = None                 # callable (dz, h, w) -> preferred block width; None: the whole width
    _forced_cols = 0

    def paint_block_cols(self, shape):
        dz, h, w = shape
        return int(self.block_rule(dz, h, w)) if self.block_rule else 1 << 20

    def paint_set_block_cols(self, cols):
        self._forced_cols = int(cols)

    def paint_open(self, ridge_own, max_rsq):
        self._ses = lt_oracle.SepPaintSession(_np(ridge_own).astype(np.uint32))
        self._out = None
        w = self._ses.w
        cols = min(self._forced_cols or self.paint_block_cols(tuple(ridge_own.shape)), w)
        self._blocks = [(x0, min(cols, w - x0)) for x0 in range(0, w, cols)]
        self._imports = None
        return len(self._blocks)

    def paint_xslab_cols(self, xs):
        return self._blocks[xs][1] * self._ses.h

    def _block_columns(self, xs):
        """whole-width column index (y * w + x) of the block's columns, in the block's order (y * xw + lx)"""
        x0, xw = self._blocks[xs]
        return (np.arange(self._ses.h)[:, None] * self._ses.w + x0 + np.arange(xw)[None, :]).reshape(-1)

    def paint_lists(self, xs):
        pass

    def boundary_buffer(self, ncol, kcap):
        return torch.zeros(ncol * (2 * kcap + 1), dtype=torch.int32)

    @staticmethod
    def _split(buf, ncol, kcap):
        a = _np(buf).view(np.uint32)
        return a[2 * ncol * kcap:], a[:2 * ncol * kcap]

    def paint_export(self, xs, side, gap, nplanes, buf, ncol, kcap):
        assert ncol == self.paint_xslab_cols(xs)
        cnt, items = self._ses.export(side, gap, nplanes, kcap)            # whole width; keep this block's columns
        sel = self._block_columns(xs)
        c, it = self._split(buf, ncol, kcap)
        c[:] = cnt[sel]
        it[:] = items[sel].reshape(-1)

    def paint_sweep(self, xs, lo, hi, ncol, kcap):
        # the model sweeps the whole width at once: collect the blocks' imports, sweep after the last block
        assert ncol == self.paint_xslab_cols(xs)
        ncol_all = self._ses.w * self._ses.h
        if self._imports is None:
            mk = lambda n: [(np.zeros(ncol_all, np.uint32), np.zeros((ncol_all, kcap, 2), np.uint32)) for _ in range(n)]
            self._imports = (mk(len(lo)), mk(len(hi)))
        sel = self._block_columns(xs)
        for bufs, full in ((lo, self._imports[0]), (hi, self._imports[1])):
            assert len(bufs) == len(full)
            for b, (cnt, items) in zip(bufs, full):
                c, it = self._split(b, ncol, kcap)
                cnt[sel] = c
                items[sel] = it.reshape(-1, kcap, 2)
        if xs == len(self._blocks) - 1:
            conv = lambda full: [(c, np.ascontiguousarray(it.reshape(-1))) for c, it in full]
            self._out = self._ses.sweep(conv(self._imports[0]), conv(self._imports[1]), kcap)

    def paint_close(self):
        flags = self._ses.close()
        return torch.from_numpy(self._out.astype(np.int32)), int(flags != 0)

    def plane_weights(self, ridge):
        """deliberately lopsided, so the painter's split differs from the I/O split in the tests"""
        r = _np(ridge).astype(np.float64)
        return (1.0 + 50.0 * np.sqrt(r).sum(axis=(1, 2))).tolist()

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
