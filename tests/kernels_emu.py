"""Loads tests/_build/libkernels_emu.so: the product's kernel sources (edt.cu, ridge.cu, paint_sep.cu) compiled by
g++ against a CPU emulation of CUDA (tests/cuda_emu/).  TEST ONLY -- it checks kernel logic without a GPU."""
import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(_HERE, "cuda_emu"))
import emu_build as _build  # noqa: E402

_lib = None


def build(force=False):
    return _build.build(force)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def set_guard_mode(mode):
    """0: an inaccessible page directly behind every array the kernels see, 1: directly in front"""
    lib().emu_set_guard_mode(int(mode))


def edt_sq(vol_u8, inverse=False, threshold=128):
    v = np.ascontiguousarray(vol_u8, np.uint8)
    d, h, w = v.shape
    sq = np.empty(v.shape, np.uint32)
    mv = C.c_uint32(0)
    lib().emu_edt(_p(v), w, h, d, int(inverse), int(threshold), _p(sq), C.byref(mv))
    return sq, int(mv.value)


def ridge(sq_u32):
    s = np.ascontiguousarray(sq_u32, np.uint32)
    d, h, w = s.shape
    out = np.empty(s.shape, np.uint32)
    cnt = C.c_ulonglong(0)
    rc = lib().emu_ridge(_p(s), w, h, d, int(s.max()), _p(out), C.byref(cnt))
    assert rc > 0
    return out, int(cnt.value)


def paint(ridge_u32, variant=0, xslab_cols=0, cap=0, fmax=0):
    r = np.ascontiguousarray(ridge_u32, np.uint32)
    d, h, w = r.shape
    out = np.empty(r.shape, np.uint32)
    flags, redo = C.c_uint32(0), C.c_ulonglong(0)
    rc = lib().emu_paint(_p(r), w, h, d, int(variant), int(xslab_cols), int(cap), int(fmax), _p(out), C.byref(flags), C.byref(redo))
    assert rc > 0
    return out, int(flags.value), int(redo.value)


def paint_slabs(ridge_u32, cuts, rmax, kcap=32, variant=0, xslab_cols=0):
    """the stepwise painter on z-slabs cut at `cuts`, staircases handed over in memory; (painted r^2, flags)"""
    r = np.ascontiguousarray(ridge_u32, np.uint32)
    d, h, w = r.shape
    out = np.empty(r.shape, np.uint32)
    c = np.ascontiguousarray(cuts, np.int32)
    rc = lib().emu_paint_slabs(_p(r), w, h, d, _p(c), len(cuts) - 1, int(rmax), int(kcap), int(variant), int(xslab_cols), _p(out))
    return out, rc


def finish(rsq_u32, original_u8, inverse=False, pixel_width=1.0, z_lo=0, z_hi=None):
    """2 sqrt + cleanup + mask (when original_u8 is given) + NaN / calibration + statistics of planes [z_lo, z_hi):
    (map float32, dict(sum, sum2, min, max, count))"""
    r = np.ascontiguousarray(rsq_u32, np.uint32)
    d, h, w = r.shape
    z_hi = d if z_hi is None else z_hi
    o = None if original_u8 is None else np.ascontiguousarray(original_u8, np.uint8)
    out = np.empty((z_hi - z_lo, h, w), np.float32)
    res = (C.c_double * 5)()
    L = lib()
    L.emu_finish.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                             C.c_void_p, C.POINTER(C.c_double)]
    rc = L.emu_finish(_p(r), None if o is None else _p(o), w, h, d, z_lo, z_hi, int(inverse), 1, float(pixel_width), _p(out), res)
    assert rc > 0
    return out, dict(sum=res[0], sum2=res[1], min=res[2], max=res[3], count=int(res[4]))


def paint_disc(ridge_u32, z_lo=0, z_hi=None, chunk_planes=0, items_per_voxel=3):
    """fronts along z + discs in the plane (paint path 4) over planes [z_lo, z_hi): (painted r^2 of those planes, flags)"""
    r = np.ascontiguousarray(ridge_u32, np.uint32)
    d, h, w = r.shape
    z_hi = d if z_hi is None else z_hi
    out = np.empty((z_hi - z_lo, h, w), np.uint32)
    flags = C.c_uint32(0)
    rc = lib().emu_paint_disc(_p(r), w, h, d, int(z_lo), int(z_hi), int(chunk_planes), int(items_per_voxel), _p(out), C.byref(flags))
    assert rc > 0, rc
    return out, int(flags.value)


def paint_scatter(ridge_u32):
    """the scatter fallback painter (paint path 2)"""
    r = np.ascontiguousarray(ridge_u32, np.uint32)
    d, h, w = r.shape
    out = np.empty(r.shape, np.uint32)
    assert lib().emu_paint_scatter(_p(r), w, h, d, _p(out)) > 0
    return out


def phantom(shapes_i32, w, h, d, z0=0, dz=None):
    s = np.ascontiguousarray(shapes_i32, np.int32)
    dz = d - z0 if dz is None else dz
    out = np.empty((dz, h, w), np.uint8)
    assert lib().emu_phantom(_p(s), int(s.shape[0]), w, h, d, z0, dz, _p(out)) > 0
    return out


class _FinishStats:
    def __init__(self, res):
        self.sum, self.sum2, self.min, self.max, self.count = res[0], res[1], res[2], res[3], int(res[4])


class EmuContext:
    """The subset of bonej2_b200._native.Context that bonej2_b200.sharded.CudaStages calls, served by the kernels on the
    CPU emulation: CudaStages(EmuContext(), torch.device("cpu")) runs the product's multi-GPU stage code on CPU tensors
    (their data_ptr() are host addresses).  One process, one painter session at a time, like a bjt_ctx."""

    def __init__(self, block_cols_rule=None):
        self._L = lib()
        self._ses = None
        self._forced = 0
        self._rule = block_cols_rule          # tests: override of the library's column-block rule (w, h, d) -> width
        L = self._L
        L.emu_paint_open.restype = C.c_void_p
        L.emu_paint_open.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_longlong]
        for f in (L.emu_paint_nxslabs, L.emu_paint_close):
            f.argtypes = [C.c_void_p]
        L.emu_paint_xslab_width.argtypes = [C.c_void_p, C.c_int]
        L.emu_paint_lists.argtypes = [C.c_void_p, C.c_int]
        L.emu_paint_export.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.emu_paint_sweep.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.emu_dev_finish_range.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                           C.c_void_p, C.POINTER(C.c_double)]
        for f in (L.emu_dev_edt_x, L.emu_dev_edt_y, L.emu_dev_edt_z, L.emu_dev_ridge, L.emu_dev_paint):
            f.argtypes = None

    @staticmethod
    def _vp(p):
        return C.c_void_p(int(p))

    def dev_edt_x(self, d_in, w, h, d, inverse, threshold, d_gx):
        assert self._L.emu_dev_edt_x(self._vp(d_in), w, h, d, int(inverse), int(threshold), self._vp(d_gx)) > 0

    def dev_edt_y(self, d_gx, w, h, d, n_sentinel, d_gy):
        assert self._L.emu_dev_edt_y(self._vp(d_gx), w, h, d, int(n_sentinel), self._vp(d_gy)) > 0

    def dev_edt_z(self, d_gy, w, h, d, n_sentinel, d_sq):
        assert self._L.emu_dev_edt_z(self._vp(d_gy), w, h, d, int(n_sentinel), self._vp(d_sq)) > 0

    def dev_ridge(self, d_sq, w, h, d, d_ridge):
        assert self._L.emu_dev_ridge(self._vp(d_sq), w, h, d, self._vp(d_ridge)) >= 0

    def dev_paint(self, d_ridge, w, h, d, d_rsq, path=-1):
        assert self._L.emu_dev_paint(self._vp(d_ridge), w, h, d, self._vp(d_rsq)) > 0

    def dev_paint_range(self, d_ridge, w, h, d, z_lo, z_hi, max_rsq, d_rsq):
        flags = C.c_uint32(0)
        rc = self._L.emu_paint_disc(self._vp(d_ridge), w, h, d, int(z_lo), int(z_hi), 0, 24, self._vp(d_rsq), C.byref(flags))
        assert rc > 0 and flags.value == 0, (rc, flags.value)

    def dev_paint_open(self, d_ridge, w, h, d, max_rsq):
        assert self._ses is None
        if self._rule and not self._forced:      # like the library: without a set width, the slab's own choice
            self._L.emu_paint_set_block_cols(int(self._rule(w, h, d)))
        self._ses = C.c_void_p(self._L.emu_paint_open(self._vp(d_ridge), w, h, d, int(max_rsq)))
        return self._L.emu_paint_nxslabs(self._ses)

    def dev_paint_xslab_width(self, xs):
        return self._L.emu_paint_xslab_width(self._ses, xs)

    def dev_paint_lists(self, xs):
        assert self._L.emu_paint_lists(self._ses, xs) > 0

    def dev_paint_export(self, xs, side, gap, nplanes, d_count, d_items, kcap):
        assert self._L.emu_paint_export(self._ses, xs, side, gap, nplanes, self._vp(d_count), self._vp(d_items), kcap) > 0

    def dev_paint_sweep(self, xs, lo, hi, kcap, d_rsq):
        def arr(pairs, k):
            a = (C.c_void_p * max(1, len(pairs)))()
            for i, pr in enumerate(pairs):
                a[i] = int(pr[k])
            return a
        assert self._L.emu_paint_sweep(self._ses, xs, len(lo), arr(lo, 0), arr(lo, 1), len(hi), arr(hi, 0), arr(hi, 1), kcap,
                                       self._vp(d_rsq)) > 0

    def dev_paint_close(self):
        f = self._L.emu_paint_close(self._ses)
        self._ses = None
        self._L.emu_paint_set_block_cols(self._forced)
        return int(f)

    def paint_block_cols(self, w, h, d):
        return int(self._rule(w, h, d)) if self._rule else int(self._L.emu_paint_block_cols(w, h, d))

    def paint_set_block_cols(self, cols):
        self._forced = int(cols)
        self._L.emu_paint_set_block_cols(int(cols))

    def dev_finish_range(self, d_rsq, d_original, w, h, d, z_lo, z_hi, inverse, calibrate, pixel_width, d_out):
        res = (C.c_double * 5)()
        o = None if d_original is None else self._vp(d_original)
        assert self._L.emu_dev_finish_range(self._vp(d_rsq), o, w, h, d, int(z_lo), int(z_hi), int(inverse), int(calibrate),
                                            float(pixel_width), self._vp(d_out), res) > 0
        return _FinishStats(res)
