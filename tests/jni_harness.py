"""Builds and loads the JNI shim test harness (tests/jni_fake_env.cpp + bonej2_b200/csrc/jni_shim.cpp compiled against
the test double tests/jni_stub/jni.h).  Two builds: linked to the product library (the `-m gpu` tests) and linked to
libbonejthickness_emu.so, the same C-ABI on the CPU emulation of CUDA (the CPU run).  TEST ONLY: the image has no JDK,
so this is how the JNI half of the seam is kept under test; a JVM never loads this file."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_OUT = os.path.join(_HERE, "_build", "libjnishim_test.so")
_OUT_EMU = os.path.join(_HERE, "_build", "libjnishim_test_emu.so")
_SRCS = [os.path.join(_HERE, "jni_fake_env.cpp"), os.path.join(_ROOT, "bonej2_b200", "csrc", "jni_shim.cpp")]
_DEPS = _SRCS + [os.path.join(_HERE, "jni_stub", "jni.h"), os.path.join(_ROOT, "include", "bonej_thickness.h")]


def build(force: bool = False, emulated: bool = False) -> str:
    if emulated:
        from tests import emu_native
        target = emu_native.build()
        out, libname = _OUT_EMU, "bonejthickness_emu"
        rpath = "$ORIGIN"
    else:
        from bonej2_b200 import _native
        target = _native.build()
        out, libname = _OUT, "bonejthickness"
        rpath = "$ORIGIN/../../bonej2_b200"
    stale = (not os.path.exists(out)) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in _DEPS + [target])
    if force or stale:
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-Wall", "-I", os.path.join(_HERE, "jni_stub"),
                               "-o", out] + _SRCS + ["-L", os.path.dirname(target), "-l" + libname, "-Wl,-rpath," + rpath])
    return out


_libs = {}


def lib(emulated: bool = False):
    if emulated not in _libs:
        L = C.CDLL(build(emulated=emulated))
        L.jt_create.restype = C.c_longlong
        L.jt_create.argtypes = [C.c_int, C.c_char_p, C.c_int]
        L.jt_create_multi.restype = C.c_longlong
        L.jt_create_multi.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_char_p, C.c_int]
        L.jt_destroy.argtypes = [C.c_longlong]
        L.jt_destroy.restype = None
        L.jt_thickness.restype = C.c_int
        L.jt_thickness.argtypes = [C.c_longlong, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                   C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_char_p, C.c_int, C.c_int, C.c_int]
        L.jt_thickness_both.restype = C.c_int
        L.jt_thickness_both.argtypes = [C.c_longlong, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                        C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_char_p, C.c_int]
        _libs[emulated] = L
    return _libs[emulated]
