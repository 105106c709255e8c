"""Builds and loads the JNI shim test harness (tests/jni_fake_env.cpp + bonej2_b200/csrc/jni_shim.cpp compiled against
the test double tests/jni_stub/jni.h, linked to the product library).  TEST ONLY: the image has no JDK, so this is how
the JNI half of the seam is kept under test; a JVM never loads this file."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_OUT = os.path.join(_HERE, "_build", "libjnishim_test.so")
_SRCS = [os.path.join(_HERE, "jni_fake_env.cpp"), os.path.join(_ROOT, "bonej2_b200", "csrc", "jni_shim.cpp")]
_DEPS = _SRCS + [os.path.join(_HERE, "jni_stub", "jni.h"), os.path.join(_ROOT, "include", "bonej_thickness.h")]


def build(force: bool = False) -> str:
    from bonej2_b200 import _native
    _native.build()
    stale = (not os.path.exists(_OUT)) or any(os.path.getmtime(s) > os.path.getmtime(_OUT) for s in _DEPS)
    if force or stale:
        os.makedirs(os.path.dirname(_OUT), exist_ok=True)
        libdir = os.path.dirname(_native.SO_PATH)
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-Wall", "-I", os.path.join(_HERE, "jni_stub"),
                               "-o", _OUT] + _SRCS + ["-L", libdir, "-lbonejthickness", "-Wl,-rpath,$ORIGIN/../../bonej2_b200"])
    return _OUT


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.jt_create.restype = C.c_longlong
        L.jt_create.argtypes = [C.c_int, C.c_char_p, C.c_int]
        L.jt_destroy.argtypes = [C.c_longlong]
        L.jt_destroy.restype = None
        L.jt_thickness.restype = C.c_int
        L.jt_thickness.argtypes = [C.c_longlong, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                   C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_char_p, C.c_int]
        _lib = L
    return _lib
