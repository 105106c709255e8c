"""ctypes binding of oracle/lt_oracle.c (the CPU restatement of sc.fiji LocalThickness_ +
ij StackStatistics as driven by ThicknessWrapper.java:123-222).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs, never by the product."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblt_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile)."""
    srcs = [os.path.join(_HERE, f) for f in ("lt_oracle.c", "phantom_cpu.c", "edt_exact.c", "sep_paint_model.cpp", "Makefile",
                                              "../bonej2_b200/csrc/paint_sep_core.cuh")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _SO


class _Stats(C.Structure):
    _fields_ = [("count", C.c_int64), ("mean", C.c_double), ("sd", C.c_double),
                ("min", C.c_double), ("max", C.c_double), ("sum", C.c_double), ("sum2", C.c_double)]


@dataclass
class Stats:
    count: int
    mean: float
    sd: float
    min: float
    max: float
    sum: float
    sum2: float


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def default_threads() -> int:
    return os.cpu_count() or 1


def _vol(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    assert a.ndim == 3, "volumes are [z][y][x]"
    return a


def edt(vol_u8, inverse=False, threshold=128, nthreads=None):
    """EDT_S1D: returns (float32 distance map, uint32 squared map); background -> 0."""
    v = _vol(vol_u8, np.uint8)
    d, h, w = v.shape
    dist = np.empty(v.shape, np.float32)
    sq = np.empty(v.shape, np.uint32)
    lib().lto_edt(_p(v, C.c_uint8), w, h, d, int(inverse), int(threshold), _p(dist, C.c_float),
                  _p(sq, C.c_uint32), nthreads or default_threads())
    return dist, sq


def edt_exact_sq(vol_u8, inverse=False, threshold=128, nthreads=None):
    """the squared map of EDT_S1D by an independent O(N) algorithm (oracle/edt_exact.c): the checker for volumes the
    brute-force restatement above cannot finish"""
    v = _vol(vol_u8, np.uint8)
    d, h, w = v.shape
    sq = np.empty(v.shape, np.uint32)
    lib().edt_exact_sq(_p(v, C.c_uint8), w, h, d, int(inverse), int(threshold), _p(sq, C.c_uint32), nthreads or default_threads())
    return sq


def ridge(dist_f32):
    """Distance_Ridge: float distance at ridge points, 0 elsewhere."""
    s = _vol(dist_f32, np.float32)
    d, h, w = s.shape
    out = np.empty(s.shape, np.float32)
    lib().lto_ridge(_p(s, C.c_float), w, h, d, _p(out, C.c_float))
    return out


def ridge_template(dist_sq_values):
    v = np.ascontiguousarray(dist_sq_values, np.int32)
    t = [np.empty(v.shape, np.int32) for _ in range(3)]
    lib().lto_ridge_template(_p(v, C.c_int32), int(v.size), *[_p(x, C.c_int32) for x in t])
    return t


def paint(ridge_f32, nthreads=None):
    """Local_Thickness_Parallel: returns (float32 2*sqrt map, uint32 painted r^2)."""
    s = _vol(ridge_f32, np.float32)
    d, h, w = s.shape
    out = np.empty(s.shape, np.float32)
    rsq = np.empty(s.shape, np.uint32)
    lib().lto_paint(_p(s, C.c_float), w, h, d, _p(out, C.c_float), _p(rsq, C.c_uint32),
                    nthreads or default_threads())
    return out, rsq


def cleanup(lt_f32):
    s = _vol(lt_f32, np.float32)
    d, h, w = s.shape
    out = np.empty(s.shape, np.float32)
    lib().lto_cleanup(_p(s, C.c_float), w, h, d, _p(out, C.c_float))
    return out


def mask(original_u8, map_f32, inverse=False):
    o = _vol(original_u8, np.uint8)
    m = np.array(map_f32, dtype=np.float32, order="C", copy=True)
    lib().lto_mask(_p(o, C.c_uint8), C.c_size_t(o.size), int(inverse), _p(m, C.c_float))
    return m


def calibrate(map_f32, pixel_width=1.0):
    m = np.array(map_f32, dtype=np.float32, order="C", copy=True)
    lib().lto_calibrate(_p(m, C.c_float), C.c_size_t(m.size), C.c_double(pixel_width))
    return m


def stats(map_f32) -> Stats:
    m = np.ascontiguousarray(map_f32, np.float32)
    st = _Stats()
    lib().lto_stats(_p(m, C.c_float), C.c_size_t(m.size), C.byref(st))
    return Stats(*(getattr(st, f) for f, _ in _Stats._fields_))


def process_image(vol_u8, inverse=False, mask=True, threshold=128, pixel_width=1.0, calibrate=True,
                  taps=False, nthreads=None):
    """LocalThicknessWrapper.processImage.  Returns None for a non-binary image (as upstream).
    Otherwise dict(map=float32 [z][y][x], stats=Stats[, sq, ridge, rsq])."""
    v = _vol(vol_u8, np.uint8)
    d, h, w = v.shape
    out = np.empty(v.shape, np.float32)
    sq = np.empty(v.shape, np.uint32) if taps else None
    rg = np.empty(v.shape, np.float32) if taps else None
    rsq = np.empty(v.shape, np.uint32) if taps else None
    st = _Stats()
    f = lib().lto_process_image
    rc = f(_p(v, C.c_uint8), w, h, d, int(inverse), int(mask), int(threshold), C.c_double(pixel_width),
           int(calibrate), _p(out, C.c_float), _p(sq, C.c_uint32), _p(rg, C.c_float), _p(rsq, C.c_uint32),
           C.byref(st), nthreads or default_threads())
    if rc != 0:
        return None
    res = dict(map=out, stats=Stats(*(getattr(st, f) for f, _ in _Stats._fields_)))
    if taps:
        res.update(sq=sq, ridge=rg, rsq=rsq)
    return res


def phantom_raster(shapes_i32, w, h, d):
    s = np.ascontiguousarray(shapes_i32, np.int32)
    assert s.ndim == 2 and s.shape[1] == 16
    out = np.empty((d, h, w), np.uint8)
    lib().lto_phantom_raster(_p(s, C.c_int32), int(s.shape[0]), w, h, d, _p(out, C.c_uint8))
    return out


class _SpmStats(C.Structure):
    _fields_ = [("overflow_act", C.c_long), ("overflow_front", C.c_long), ("max_active", C.c_long),
                ("max_front", C.c_long), ("max_stair", C.c_long), ("items", (C.c_long * 2) * 2),
                ("sum_active", C.c_double * 3), ("steps", C.c_long * 3)]


def sep_paint_model(ridge_u32):
    """CPU model of the separable painter (oracle/sep_paint_model.cpp, running the per-line code of
    bonej2_b200/csrc/paint_sep_core.cuh): (painted r^2, stats dict)."""
    r = _vol(ridge_u32, np.uint32)
    d, h, w = r.shape
    out = np.empty(r.shape, np.uint32)
    st = _SpmStats()
    rc = lib().spm_paint(_p(r, C.c_uint32), w, h, d, _p(out, C.c_uint32), C.byref(st))
    info = dict(rc=rc, overflow_act=st.overflow_act, overflow_front=st.overflow_front, max_active=st.max_active,
                max_front=st.max_front, max_stair=st.max_stair,
                items_per_voxel=[[st.items[s][k] / r.size for k in range(2)] for s in range(2)],
                mean_active=[st.sum_active[s] / max(1, st.steps[s]) for s in range(3)])
    return out, info


def sep_patch_map(na, nb, threads_per_block=128):
    """thread -> line map of the painter kernels (sep_patch_line in paint_sep_core.cuh) for a plane of na x nb
    lines and a grid of whole blocks: int64 array, -1 = idle thread."""
    L = lib()
    L.spm_patch_map.restype = C.c_longlong
    L.spm_patch_map.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong), C.c_longlong]
    cap = ((na + 7) // 8) * ((nb + 3) // 4) * 32 + threads_per_block
    lines = np.empty(cap, np.int64)
    n = L.spm_patch_map(na, nb, threads_per_block, _p(lines, C.c_longlong), cap)
    assert n >= 0
    return lines[:n]


class SepPaintSession:
    """The CPU model's painter in steps (oracle/sep_paint_model.cpp spm_open .. spm_close): same calls and
    buffer layout as bjt_dev_paint_open .. _close, on numpy arrays.  TEST ONLY."""

    def __init__(self, ridge_u32):
        r = _vol(ridge_u32, np.uint32)
        self.d, self.h, self.w = r.shape
        L = lib()
        L.spm_open.restype = C.c_void_p
        L.spm_close.restype = C.c_long
        self._h = C.c_void_p(L.spm_open(_p(r, C.c_uint32), self.w, self.h, self.d))

    def export(self, side, gap, nplanes, kcap):
        ncol = self.w * self.h
        cnt = np.zeros(ncol, np.uint32)
        items = np.zeros((ncol, kcap, 2), np.uint32)
        lib().spm_export(self._h, int(side), int(gap), int(nplanes), _p(cnt, C.c_uint32), _p(items, C.c_uint32), int(kcap))
        return cnt, items

    def sweep(self, lo, hi, kcap):
        """lo / hi: lists of (cnt, items) imported from below / above"""
        def arrs(pairs):
            PT = C.POINTER(C.c_uint32)
            a = (PT * max(1, len(pairs)))()
            b = (PT * max(1, len(pairs)))()
            for i, (c, it) in enumerate(pairs):
                a[i] = _p(c, C.c_uint32)
                b[i] = _p(it, C.c_uint32)
            return a, b
        out = np.zeros((self.d, self.h, self.w), np.uint32)
        lc, li = arrs(lo)
        hc, hi_ = arrs(hi)
        lib().spm_sweep(self._h, len(lo), lc, li, len(hi), hc, hi_, int(kcap), _p(out, C.c_uint32))
        return out

    def close(self) -> int:
        return int(lib().spm_close(self._h))
