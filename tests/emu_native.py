"""TEST ONLY: binds bonej2_b200._native to libbonejthickness_emu.so -- the library's whole C-ABI (api.cu, the kernels,
thickness_wrapper.cpp) compiled against the CPU emulation of CUDA in tests/cuda_emu/ -- for the duration of a `with`
block, so the tests that drive the C-ABI on a GPU can also run, at small sizes, where there is none.  Nothing in the
product loads this library: bonej2_b200._native.lib() only ever opens the nvcc-built libbonejthickness.so and fails
without it."""
import contextlib
import ctypes as C
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(_HERE, "cuda_emu"))
import emu_build as _build  # noqa: E402

from bonej2_b200 import _native  # noqa: E402


def build(force=False):
    return _build.build_api(force)


@contextlib.contextmanager
def emulated_library():
    L = C.CDLL(build())
    for name, (args, res) in _native._SIGS.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = res
    from bonej2_b200 import wrapper
    saved = (_native._lib, _native.SO_PATH, wrapper._wlib)
    _native._lib = L
    _native.SO_PATH, wrapper._wlib = _build.API_OUT, None     # the C++ ThicknessWrapper mirror lives in the same library
    wrapper._wl()
    _native.SO_PATH = saved[1]
    try:
        yield L
    finally:
        _native._lib, _native.SO_PATH, wrapper._wlib = saved
