"""ctypes binding of libbonejthickness.so (include/bonej_thickness.h) -- the same C-ABI a JNI shim
binds for ThicknessWrapper.createMap (ThicknessWrapper.java:190-196).  No CPU fallback: if the
library or a CUDA device is missing, every call raises."""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
SO_PATH = os.path.join(_HERE, "libbonejthickness.so")
CSRC = os.path.join(_HERE, "csrc")
HEADER = os.path.join(_ROOT, "include", "bonej_thickness.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

BJT_OK, BJT_E_CUDA, BJT_E_ARG, BJT_E_NOT_BINARY, BJT_E_INTERNAL = 0, -1, -2, -3, -4


class ThicknessError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libbonejthickness error {code}: {msg}")
        self.code = code


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into the in-tree shared library (nvcc cross-compiles
    without a GPU)."""
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(_ROOT, "include", "*.h*"))
    stale = (not os.path.exists(SO_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(SO_PATH) for s in deps)
    if force or stale:
        nvcc = os.environ.get("NVCC", "nvcc")
        extra = os.environ.get("BJT_NVCC_EXTRA", "").split()        # tuning builds (tools/tune_painter.sh), e.g. -DBJT_S_SM=10
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", SO_PATH] + sources()
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return SO_PATH


class Stats(C.Structure):
    _fields_ = [("count", C.c_int64), ("mean", C.c_double), ("sd", C.c_double), ("min", C.c_double),
                ("max", C.c_double), ("sum", C.c_double), ("sum2", C.c_double)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class Timing(C.Structure):
    _fields_ = [("ms_h2d", C.c_float), ("ms_edt_x", C.c_float), ("ms_edt_y", C.c_float), ("ms_edt_z", C.c_float),
                ("ms_ridge", C.c_float), ("ms_paint", C.c_float), ("ms_cleanup", C.c_float),
                ("ms_finish", C.c_float), ("ms_d2h", C.c_float), ("ms_total", C.c_float),
                ("n_ridge", C.c_int64), ("n_launches", C.c_int64), ("rsq_max", C.c_uint32),
                ("paint_path", C.c_int32), ("ms_paint_x", C.c_float), ("ms_paint_y", C.c_float),
                ("ms_paint_z", C.c_float), ("n_paint_y_sweeps", C.c_int32)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


_lib = None

_VP, _I, _D = C.c_void_p, C.c_int, C.c_double
_SIGS = {
    "bjt_create": ([C.POINTER(_VP), _I], _I),
    "bjt_create_multi": ([C.POINTER(_VP), C.POINTER(_I), _I], _I),
    "bjt_device_count": ([_VP], _I),
    "bjt_destroy": ([_VP], None),
    "bjt_last_error": ([_VP], C.c_char_p),
    "bjt_set_stream": ([_VP, _VP], _I),
    "bjt_release_workspace": ([_VP], _I),
    "bjt_device_info": ([_VP, C.c_char_p, _I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)], _I),
    "bjt_set_paint_path": ([_VP, _I], _I),
    "bjt_launch_count": ([_VP], C.c_int64),
    "bjt_find_ridge_points_u8": ([_VP, _VP, _I, _I, _I, _D, _VP, C.c_int64, _VP, _VP], _I),
    "bjt_label_stats_f32": ([_VP, _VP, _VP, C.c_size_t, _I, _VP, _VP], _I),
    "bjt_dev_label_stats_f32": ([_VP, _VP, _VP, C.c_size_t, _I, _VP, _VP], _I),
    "bjt_slice_stats_f32": ([_VP, _VP, _I, _I, _I, _I, _I, _I, _I, _VP, _VP], _I),
    "bjt_dev_slice_stats_f32": ([_VP, _VP, _I, _I, _I, _I, _I, _I, _I, _VP, _VP], _I),
    "bjt_dev_paint_range": ([_VP, _VP, _I, _I, _I, _I, _I, C.c_int64, _VP], _I),
    "bjt_dev_paint_open": ([_VP, _VP, _I, _I, _I, C.c_int64], _I),
    "bjt_dev_paint_nxslabs": ([_VP], _I),
    "bjt_dev_paint_xslab_width": ([_VP, _I], _I),
    "bjt_dev_paint_lists": ([_VP, _I], _I),
    "bjt_dev_paint_export": ([_VP, _I, _I, _I, _I, _VP, _VP, _I], _I),
    "bjt_dev_paint_sweep": ([_VP, _I, _I, _VP, _VP, _I, _VP, _VP, _I, _VP], _I),
    "bjt_dev_paint_close": ([_VP, _VP], _I),
    "bjt_debug_set_paint_limits": ([_VP, _I, _I], _I),
    "bjt_debug_paint_redo_lines": ([_VP], C.c_int64),
    "bjt_debug_isqrt_check": ([_VP, C.POINTER(C.c_int64)], _I),
    "bjt_debug_set_xslab_cols": ([_I], _I),
    "bjt_paint_block_cols": ([_I, _I, _I], _I),
    "bjt_paint_set_block_cols": ([_I], _I),
    "bjt_host_alloc": ([C.c_size_t], _VP),
    "bjt_host_free": ([_VP], None),
    "bjt_thickness_u8": ([_VP, _VP, _I, _I, _I, _I, _I, _I, _D, _VP, C.POINTER(Stats), C.POINTER(Timing)], _I),
    "bjt_thickness_u8_slices": ([_VP, _VP, _I, _I, _I, _I, _I, _I, _D, _VP, C.POINTER(Stats), C.POINTER(Timing)], _I),
    "bjt_thickness_both_u8": ([_VP, _VP, _I, _I, _I, _I, _I, _D, _VP, _VP, C.POINTER(Stats), C.POINTER(Stats),
                               C.POINTER(Timing), C.POINTER(Timing)], _I),
    "bjt_dev_thickness_u8": ([_VP, _VP, _I, _I, _I, _I, _I, _I, _D, _VP, C.POINTER(Stats), C.POINTER(Timing)], _I),
    "bjt_edt_sq_u8": ([_VP, _VP, _I, _I, _I, _I, _I, _VP], _I),
    "bjt_ridge_u32": ([_VP, _VP, _I, _I, _I, _VP], _I),
    "bjt_ridge_template": ([_VP, _VP, _I, _VP, _VP, _VP], _I),
    "bjt_paint_rsq": ([_VP, _VP, _I, _I, _I, _I, _VP], _I),
    "bjt_cleanup_f32": ([_VP, _VP, _VP, _I, _I, _I, _I, _I, _D, _VP, C.POINTER(Stats)], _I),
    "bjt_stats_f32": ([_VP, _VP, C.c_size_t, C.POINTER(Stats)], _I),
    "bjt_dev_edt_x": ([_VP, _VP, _I, _I, _I, _I, _I, _VP], _I),
    "bjt_dev_edt_y": ([_VP, _VP, _I, _I, _I, _I, _VP], _I),
    "bjt_dev_edt_z": ([_VP, _VP, _I, _I, _I, _I, _VP], _I),
    "bjt_dev_ridge": ([_VP, _VP, _I, _I, _I, _VP], _I),
    "bjt_dev_paint": ([_VP, _VP, _I, _I, _I, _I, _VP], _I),
    "bjt_dev_finish": ([_VP, _VP, _VP, _I, _I, _I, _I, _I, _D, _VP, C.POINTER(Stats)], _I),
    "bjt_dev_finish_range": ([_VP, _VP, _VP, _I, _I, _I, _I, _I, _I, _I, _D, _VP, C.POINTER(Stats)], _I),
    "bjt_dev_alloc": ([_VP, C.c_char_p, C.c_size_t, C.POINTER(_VP)], _I),
    "bjt_ipc_export": ([_VP, _VP, C.c_char_p], _I),
    "bjt_ipc_import": ([_VP, C.c_char_p, C.POINTER(_VP)], _I),
    "bjt_ipc_close": ([_VP, _VP], _I),
    "bjt_dev_copy2d": ([_VP, _VP, C.c_size_t, _VP, C.c_size_t, C.c_size_t, C.c_size_t], _I),
    "bjt_dev_copy2d_batch": ([_VP, _I, _VP, _VP, _VP, _VP, _VP, _VP], _I),
    "bjt_phantom_u8": ([_VP, _VP, _I, _I, _I, _I, _VP], _I),
    "bjt_dev_phantom_u8": ([_VP, _VP, _I, _I, _I, _I, _I, _I, _VP], _I),
}


def exported_symbols():
    return sorted(_SIGS)


def lib():
    """Load the library (RuntimeError if it has not been built -- there is nothing to fall back to)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the Thickness path has no CPU fallback)")
        L = C.CDLL(SO_PATH)
        for name, (args, res) in _SIGS.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = res
        _lib = L
    return _lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return int(a)       # raw address (device pointer or pinned host pointer)


class Context:
    """One bjt_ctx: a CUDA device, a stream and the cached device workspace."""

    def __init__(self, device: int = -1, devices=None):
        """device: one GPU (bjt_create).  devices: a sequence of GPUs -- the thickness calls then cut the stack into
        z-slabs over them inside the library (bjt_create_multi); a GPU may be named more than once."""
        self._h = _VP()
        if devices is not None:
            arr = (C.c_int * len(devices))(*[int(x) for x in devices])
            rc = lib().bjt_create_multi(C.byref(self._h), arr, len(devices))
        else:
            rc = lib().bjt_create(C.byref(self._h), int(device))
        if rc != BJT_OK:
            raise ThicknessError(rc, lib().bjt_last_error(None).decode())

    def device_count(self) -> int:
        return int(lib().bjt_device_count(self._h))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().bjt_destroy(self._h)
            self._h = _VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc):
        if rc != BJT_OK:
            raise ThicknessError(rc, lib().bjt_last_error(self._h).decode())

    # -- configuration
    def set_stream(self, cuda_stream: int):
        self._ck(lib().bjt_set_stream(self._h, cuda_stream))

    def set_paint_path(self, path: int):
        self._ck(lib().bjt_set_paint_path(self._h, int(path)))

    def find_ridge_points(self, vol_u8, fraction=0.6, cap=1 << 20):
        """FindRidgePoints: ([n][3] seed points (x, y, z), number found, largest ridge value)"""
        v = self._vol(vol_u8, np.uint8)
        d, h, w = v.shape
        out = np.zeros((cap, 3), np.float64)
        n = C.c_int64(0)
        mx = C.c_float(0)
        self._ck(lib().bjt_find_ridge_points_u8(self._h, _ptr(v), w, h, d, float(fraction), _ptr(out), cap, C.byref(n), C.byref(mx)))
        return out[:min(cap, n.value)].copy(), int(n.value), float(mx.value)

    # ---- consumers of the map ----
    def label_stats(self, map_f32, labels_i32, sizes_i64):
        """ParticleAnalysis.getMeanStdDev: [n_labels][3] = mean, sd, max per particle label"""
        m = np.ascontiguousarray(map_f32, np.float32)
        lab = np.ascontiguousarray(labels_i32, np.int32)
        sz = np.ascontiguousarray(sizes_i64, np.int64)
        out = np.zeros((sz.size, 3), np.float64)
        self._ck(lib().bjt_label_stats_f32(self._h, _ptr(m), _ptr(lab), m.size, sz.size, _ptr(sz), _ptr(out)))
        return out

    def slice_stats(self, map_f32, roi=None):
        """SliceGeometry: ([d][3] = mean, max, sd per slice inside roi = (x, y, w, h), counts)"""
        m = self._vol(map_f32, np.float32)
        d, h, w = m.shape
        rx, ry, rw, rh = roi if roi is not None else (0, 0, w, h)
        out = np.zeros((d, 3), np.float64)
        cnt = np.zeros(d, np.int64)
        self._ck(lib().bjt_slice_stats_f32(self._h, _ptr(m), w, h, d, rx, ry, rw, rh, _ptr(out), _ptr(cnt)))
        return out, cnt

    # ---- stepwise painter of a z-slab (see include/bonej_thickness.h) ----
    def dev_paint_open(self, d_ridge, w, h, d, max_rsq):
        self._ck(lib().bjt_dev_paint_open(self._h, d_ridge, w, h, d, int(max_rsq)))
        return lib().bjt_dev_paint_nxslabs(self._h)

    def dev_paint_xslab_width(self, xs):
        return lib().bjt_dev_paint_xslab_width(self._h, xs)

    def dev_paint_lists(self, xs):
        self._ck(lib().bjt_dev_paint_lists(self._h, xs))

    def dev_paint_export(self, xs, side, gap, nplanes, d_count, d_items, kcap):
        self._ck(lib().bjt_dev_paint_export(self._h, xs, side, gap, nplanes, d_count, d_items, kcap))

    def dev_paint_sweep(self, xs, lo, hi, kcap, d_rsq):
        """lo / hi: lists of (count_ptr, items_ptr) device pointers of the staircases imported from below / above"""
        def arr(pairs, k):
            a = (C.c_void_p * max(1, len(pairs)))()
            for i, pr in enumerate(pairs):
                a[i] = pr[k]
            return a
        self._ck(lib().bjt_dev_paint_sweep(self._h, xs, len(lo), arr(lo, 0), arr(lo, 1), len(hi), arr(hi, 0), arr(hi, 1),
                                           kcap, d_rsq))

    def dev_paint_close(self) -> int:
        f = C.c_uint32(0)
        self._ck(lib().bjt_dev_paint_close(self._h, C.byref(f)))
        return int(f.value)

    def debug_set_paint_limits(self, active_cap=0, front_cap=0):
        self._ck(lib().bjt_debug_set_paint_limits(self._h, int(active_cap), int(front_cap)))

    @staticmethod
    def debug_set_xslab_cols(cols=0):
        if lib().bjt_debug_set_xslab_cols(int(cols)) != 0:
            raise ValueError("x-slab width must be a non-negative multiple of 32")

    @staticmethod
    def paint_block_cols(w, h, d) -> int:
        """column-block width the painter chooses for a w x h x d volume (not clipped to w)"""
        v = lib().bjt_paint_block_cols(int(w), int(h), int(d))
        if v <= 0:
            raise ValueError("paint_block_cols: non-positive volume size")
        return int(v)

    @staticmethod
    def paint_set_block_cols(cols=0):
        """use this column-block width (multiple of 32; 0 = the library's choice); process-wide"""
        if lib().bjt_paint_set_block_cols(int(cols)) != 0:
            raise ValueError("column-block width must be a non-negative multiple of 32")

    def debug_isqrt_check(self) -> int:
        """mismatches of the disc painter's integer square root against the exact one (must be 0)"""
        n = C.c_int64(-1)
        self._ck(lib().bjt_debug_isqrt_check(self._h, C.byref(n)))
        return int(n.value)

    def debug_paint_redo_lines(self) -> int:
        return int(lib().bjt_debug_paint_redo_lines(self._h))

    def launch_count(self) -> int:
        """kernels of this library launched through this context so far"""
        return int(lib().bjt_launch_count(self._h))

    def release_workspace(self):
        self._ck(lib().bjt_release_workspace(self._h))

    def device_info(self):
        name = C.create_string_buffer(256)
        sm, ma, mi = _I(), _I(), _I()
        self._ck(lib().bjt_device_info(self._h, name, 256, C.byref(sm), C.byref(ma), C.byref(mi)))
        return dict(name=name.value.decode(), sm_count=sm.value, cc=(ma.value, mi.value))

    # -- the seam
    @staticmethod
    def _vol(a, dtype):
        a = np.ascontiguousarray(a, dtype=dtype)
        if a.ndim != 3:
            raise ValueError("volumes are [z][y][x]")
        return a

    def thickness(self, vol_u8, inverse=False, mask=True, threshold=128, pixel_width=1.0, want_map=True, out=None):
        v = self._vol(vol_u8, np.uint8)
        d, h, w = v.shape
        if want_map and out is None:
            out = np.empty(v.shape, np.float32)
        st, tm = Stats(), Timing()
        self._ck(lib().bjt_thickness_u8(self._h, _ptr(v), w, h, d, int(inverse), int(mask), int(threshold),
                                        float(pixel_width), _ptr(out) if want_map else None, C.byref(st), C.byref(tm)))
        return (out if want_map else None), st, tm

    def thickness_slices(self, slices_u8, inverse=False, mask=True, threshold=128, pixel_width=1.0):
        """Java-style stack: a list of d separate [h][w] arrays in, a list of d float arrays out."""
        sl = [np.ascontiguousarray(s, np.uint8) for s in slices_u8]
        d = len(sl)
        h, w = sl[0].shape
        outs = [np.empty((h, w), np.float32) for _ in range(d)]
        ip = (C.c_void_p * d)(*[s.ctypes.data for s in sl])
        op = (C.c_void_p * d)(*[o.ctypes.data for o in outs])
        st, tm = Stats(), Timing()
        self._ck(lib().bjt_thickness_u8_slices(self._h, C.cast(ip, _VP), w, h, d, int(inverse), int(mask),
                                               int(threshold), float(pixel_width), C.cast(op, _VP), C.byref(st),
                                               C.byref(tm)))
        return outs, st, tm

    def thickness_both(self, vol_u8, mask=True, threshold=128, pixel_width=1.0, want_maps=True, out_th=None,
                       out_sp=None):
        v = self._vol(vol_u8, np.uint8)
        d, h, w = v.shape
        if want_maps:
            out_th = np.empty(v.shape, np.float32) if out_th is None else out_th
            out_sp = np.empty(v.shape, np.float32) if out_sp is None else out_sp
        s0, s1, t0, t1 = Stats(), Stats(), Timing(), Timing()
        self._ck(lib().bjt_thickness_both_u8(self._h, _ptr(v), w, h, d, int(mask), int(threshold), float(pixel_width),
                                             _ptr(out_th) if want_maps else None, _ptr(out_sp) if want_maps else None,
                                             C.byref(s0), C.byref(s1), C.byref(t0), C.byref(t1)))
        return (out_th, s0, t0), (out_sp, s1, t1)

    def thickness_both_raw(self, in_ptr, w, h, d, mask, threshold, pixel_width, out_th_ptr, out_sp_ptr):
        """Raw host addresses (pinned buffers of the bench)."""
        s0, s1, t0, t1 = Stats(), Stats(), Timing(), Timing()
        self._ck(lib().bjt_thickness_both_u8(self._h, in_ptr, w, h, d, int(mask), int(threshold), float(pixel_width),
                                             out_th_ptr, out_sp_ptr, C.byref(s0), C.byref(s1), C.byref(t0),
                                             C.byref(t1)))
        return (s0, t0), (s1, t1)

    def dev_thickness(self, d_in, w, h, d, inverse=False, mask=True, threshold=128, pixel_width=1.0, d_out=None):
        st, tm = Stats(), Timing()
        self._ck(lib().bjt_dev_thickness_u8(self._h, _ptr(d_in), w, h, d, int(inverse), int(mask), int(threshold),
                                            float(pixel_width), _ptr(d_out), C.byref(st), C.byref(tm)))
        return st, tm

    # -- stage level, device pointers (ints); used by the z-slab sharded orchestration
    # ---- buffers other processes read (CUDA IPC), strided device copies ----
    def dev_alloc(self, name: str, nbytes: int) -> int:
        p = _VP()
        self._ck(lib().bjt_dev_alloc(self._h, name.encode(), int(nbytes), C.byref(p)))
        return int(p.value)

    def ipc_export(self, ptr: int) -> bytes:
        buf = C.create_string_buffer(64)
        self._ck(lib().bjt_ipc_export(self._h, ptr, buf))
        return buf.raw

    def ipc_import(self, handle: bytes) -> int:
        p = _VP()
        self._ck(lib().bjt_ipc_import(self._h, handle, C.byref(p)))
        return int(p.value)

    def ipc_close(self, ptr: int):
        self._ck(lib().bjt_ipc_close(self._h, ptr))

    def dev_copy2d(self, dst, dpitch, src, spitch, width, height):
        self._ck(lib().bjt_dev_copy2d(self._h, int(dst), int(dpitch), int(src), int(spitch), int(width), int(height)))

    def dev_copy2d_batch(self, blocks):
        """blocks: [(dst, dpitch, src, spitch, width, height)] -- all of them in one launch"""
        n = len(blocks)
        cols = list(zip(*blocks)) if n else [[]] * 6
        arr = [(C.c_void_p * max(n, 1))(*cols[0]), (C.c_size_t * max(n, 1))(*cols[1]), (C.c_void_p * max(n, 1))(*cols[2]),
               (C.c_size_t * max(n, 1))(*cols[3]), (C.c_size_t * max(n, 1))(*cols[4]), (C.c_size_t * max(n, 1))(*cols[5])]
        self._ck(lib().bjt_dev_copy2d_batch(self._h, n, *arr))

    def dev_edt_x(self, d_in, w, h, d, inverse, threshold, d_gx):
        self._ck(lib().bjt_dev_edt_x(self._h, d_in, w, h, d, int(inverse), int(threshold), d_gx))

    def dev_edt_y(self, d_gx, w, h, d, n_sentinel, d_gy):
        self._ck(lib().bjt_dev_edt_y(self._h, d_gx, w, h, d, int(n_sentinel), d_gy))

    def dev_edt_z(self, d_gy, w, h, d, n_sentinel, d_sq):
        self._ck(lib().bjt_dev_edt_z(self._h, d_gy, w, h, d, int(n_sentinel), d_sq))

    def dev_ridge(self, d_sq, w, h, d, d_ridge):
        self._ck(lib().bjt_dev_ridge(self._h, d_sq, w, h, d, d_ridge))

    def dev_paint(self, d_ridge, w, h, d, d_rsq, path=-1):
        self._ck(lib().bjt_dev_paint(self._h, d_ridge, w, h, d, int(path), d_rsq))

    def dev_paint_range(self, d_ridge, w, h, d, z_lo, z_hi, max_rsq, d_rsq):
        self._ck(lib().bjt_dev_paint_range(self._h, d_ridge, w, h, d, int(z_lo), int(z_hi), int(max_rsq), d_rsq))

    def dev_finish_range(self, d_rsq, d_original, w, h, d, z_lo, z_hi, inverse, calibrate, pixel_width, d_out):
        st = Stats()
        self._ck(lib().bjt_dev_finish_range(self._h, d_rsq, d_original, w, h, d, int(z_lo), int(z_hi), int(inverse),
                                            int(calibrate), float(pixel_width), d_out, C.byref(st)))
        return st

    # -- stage level (host buffers)
    def edt_sq(self, vol_u8, inverse=False, threshold=128):
        v = self._vol(vol_u8, np.uint8)
        d, h, w = v.shape
        out = np.empty(v.shape, np.uint32)
        self._ck(lib().bjt_edt_sq_u8(self._h, _ptr(v), w, h, d, int(inverse), int(threshold), _ptr(out)))
        return out

    def ridge(self, sq_u32):
        s = self._vol(sq_u32, np.uint32)
        d, h, w = s.shape
        out = np.empty(s.shape, np.uint32)
        self._ck(lib().bjt_ridge_u32(self._h, _ptr(s), w, h, d, _ptr(out)))
        return out

    def ridge_template(self, rsq_values):
        v = np.ascontiguousarray(rsq_values, np.int32)
        t = [np.empty(v.shape, np.int32) for _ in range(3)]
        self._ck(lib().bjt_ridge_template(self._h, _ptr(v), int(v.size), _ptr(t[0]), _ptr(t[1]), _ptr(t[2])))
        return t

    def paint_rsq(self, ridge_u32, path=-1):
        r = self._vol(ridge_u32, np.uint32)
        d, h, w = r.shape
        out = np.empty(r.shape, np.uint32)
        self._ck(lib().bjt_paint_rsq(self._h, _ptr(r), w, h, d, int(path), _ptr(out)))
        return out

    def cleanup(self, rsq_u32, original_u8=None, inverse=False, calibrate=True, pixel_width=1.0):
        r = self._vol(rsq_u32, np.uint32)
        d, h, w = r.shape
        o = None if original_u8 is None else self._vol(original_u8, np.uint8)
        out = np.empty(r.shape, np.float32)
        st = Stats()
        self._ck(lib().bjt_cleanup_f32(self._h, _ptr(r), _ptr(o), w, h, d, int(inverse), int(calibrate),
                                       float(pixel_width), _ptr(out), C.byref(st)))
        return out, st

    def stats(self, map_f32):
        m = np.ascontiguousarray(map_f32, np.float32)
        st = Stats()
        self._ck(lib().bjt_stats_f32(self._h, _ptr(m), m.size, C.byref(st)))
        return st

    # -- phantom
    def phantom(self, shapes_i32, w, h, d):
        s = np.ascontiguousarray(shapes_i32, np.int32)
        out = np.empty((d, h, w), np.uint8)
        self._ck(lib().bjt_phantom_u8(self._h, _ptr(s), int(s.shape[0]), w, h, d, _ptr(out)))
        return out

    def dev_phantom(self, shapes_i32, w, h, d, z0, dz, d_out):
        s = np.ascontiguousarray(shapes_i32, np.int32)
        self._ck(lib().bjt_dev_phantom_u8(self._h, _ptr(s), int(s.shape[0]), w, h, d, z0, dz, _ptr(d_out)))
