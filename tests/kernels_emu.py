"""Loads tests/_build/libkernels_emu.so: the product's kernel sources (edt.cu, ridge.cu, paint_sep.cu) compiled by
g++ against a CPU emulation of CUDA (tests/cuda_emu/).  TEST ONLY -- it checks kernel logic without a GPU."""
import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(_HERE, "cuda_emu"))
import build as _build  # noqa: E402

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
