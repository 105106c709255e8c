"""ctypes driver of the C++ ThicknessWrapper mirror (include/thickness_wrapper.hpp,
bonej2_b200/csrc/thickness_wrapper.cpp) so the tests can read like
Modern/wrapperPlugins/src/test/java/org/bonej/wrapperPlugins/ThicknessWrapperTest.java."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _native

X, Y, Z, CHANNEL, TIME, UNKNOWN = range(6)


class _Dataset(C.Structure):
    _fields_ = [("name", C.c_char_p), ("n_axes", C.c_int), ("axis_types", C.POINTER(C.c_int)),
                ("axis_sizes", C.POINTER(C.c_long)), ("axis_scales", C.POINTER(C.c_double)),
                ("axis_units", C.POINTER(C.c_char_p)), ("bits_per_pixel", C.c_int), ("data", C.c_void_p)]


class _Result(C.Structure):
    _fields_ = [("canceled", C.c_int), ("cancel_reason", C.c_char * 256), ("log", C.c_char * 1024),
                ("trabecular_map", C.POINTER(C.c_float)), ("separation_map", C.POINTER(C.c_float)),
                ("trabecular_max", C.c_double), ("separation_max", C.c_double), ("color_table", C.c_char * 16),
                ("n_columns", C.c_int), ("n_rows", C.c_int), ("headers", (C.c_char * 64) * 16),
                ("row_labels", (C.c_char * 128) * 16), ("values", (C.c_double * 16) * 16), ("exception", C.c_int)]


@dataclass
class Dataset:
    """what the command reads of a net.imagej.Dataset: data with first axis fastest, typed calibrated axes"""
    data: np.ndarray                      # numpy order is slowest axis first, i.e. reversed(axes)
    axes: list = field(default_factory=lambda: [X, Y, Z])
    scales: list = None
    units: list = None
    name: str = "test"


@dataclass
class Outcome:
    canceled: bool
    cancel_reason: str
    log: str
    trabecularMap: np.ndarray | None
    separationMap: np.ndarray | None
    trabecular_max: float
    separation_max: float
    color_table: str
    headers: list
    row_labels: list
    table: dict                             # header -> list of values per row
    exception: int


def shared_table_reset():
    _native.lib()
    _wl().bjw_shared_table_reset()


_wlib = None


def _wl():
    global _wlib
    if _wlib is None:
        L = C.CDLL(_native.SO_PATH)
        L.bjw_thickness_run.argtypes = [C.POINTER(_Dataset), C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.POINTER(_Result)]
        L.bjw_thickness_run.restype = C.c_int
        L.bjw_result_free.argtypes = [C.POINTER(_Result)]
        L.bjw_shared_table_reset.argtypes = []
        _wlib = L
    return _wlib


def run_thickness(dataset: Dataset | None, mapChoice="Trabecular thickness", showMaps=True, maskArtefacts=True,
                  headless=True, dialog_ok=True, device=0) -> Outcome:
    """CommandService.run(ThicknessWrapper.class, true, "inputDataset", ds, "mapChoice", ..., ...).get()"""
    L = _wl()
    res = _Result()
    keep = []
    dsp = None
    shape = None
    if dataset is not None:
        arr = np.ascontiguousarray(dataset.data)
        n = len(dataset.axes)
        sizes = list(reversed(arr.shape))
        assert len(sizes) == n, "one axis per array dimension"
        scales = dataset.scales or [1.0] * n
        units = dataset.units or [""] * n
        bits = 8 if arr.dtype == np.uint8 else arr.dtype.itemsize * 8
        d = _Dataset(dataset.name.encode(), n, (C.c_int * n)(*dataset.axes), (C.c_long * n)(*sizes),
                     (C.c_double * n)(*scales), (C.c_char_p * n)(*[u.encode() for u in units]), bits,
                     arr.ctypes.data)
        keep = [arr, d]
        dsp = C.byref(d)
        spatial = {a: s for a, s in zip(dataset.axes, sizes)}
        if all(k in spatial for k in (X, Y, Z)):
            shape = (spatial[Z], spatial[Y], spatial[X])
    rc = L.bjw_thickness_run(dsp, mapChoice.encode(), int(showMaps), int(maskArtefacts), int(headless), int(dialog_ok),
                             device, C.byref(res))
    if rc != 0:
        msg = res.cancel_reason.decode()
        L.bjw_result_free(C.byref(res))
        raise RuntimeError(msg)

    def grab(p):
        if not p or shape is None:
            return None
        return np.ctypeslib.as_array(p, shape=shape).copy()

    headers = [res.headers[c].value.decode() for c in range(res.n_columns)]
    rows = [res.row_labels[r].value.decode() for r in range(res.n_rows)]
    table = {h: [res.values[c][r] for r in range(res.n_rows)] for c, h in enumerate(headers)}
    out = Outcome(bool(res.canceled), res.cancel_reason.decode(), res.log.decode(), grab(res.trabecular_map),
                  grab(res.separation_map), res.trabecular_max, res.separation_max, res.color_table.decode(), headers,
                  rows, table, res.exception)
    L.bjw_result_free(C.byref(res))
    del keep
    return out
