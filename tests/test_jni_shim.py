"""The JNI half of the seam (bonej2_b200/csrc/jni_shim.cpp, bound by java/org/bonej/wrapperPlugins/b200/
ThicknessNative.java) driven by a fake JNIEnv, because the image has no JDK (tests/jni_harness.py).  Checks the shim's
contract from INTEGRATION.md: no Java array is ever pinned -- every slice is copied exactly once through
Get/SetArrayRegion into page-locked staging, local references are dropped slice by slice --, arguments are validated
before anything is copied (IllegalArgumentException), the input is untouched, MapStats = the library's statistics,
null for a non-binary image (LocalThicknessWrapper.processImage returns null), RuntimeException with bjt_last_error()
otherwise, thicknessBoth = mapChoice "Both" in one call, createMulti = z-slabs over several devices -- and never a CPU
fallback.  Every test runs twice: against the product library on a GPU (`-m gpu`) and against the same C-ABI on the
CPU emulation of CUDA (tests/cuda_emu) in the CPU run."""
import ctypes as C

import numpy as np
import pytest

from tests import shapes

INFO = ("reads", "writes", "live_refs", "max_live_refs", "each_once", "violations")


def _call(L, h, vol, inverse, mask, want_map=True, threshold=128, pixel_width=1.0, dims=None, broken=-1, broken_kind=0):
    d, hh, w = vol.shape if dims is None else dims
    v = np.ascontiguousarray(vol, np.uint8)
    out = np.full(v.shape, -2.0, np.float32) if want_map else None
    stats = (C.c_double * 5)()
    info = (C.c_int * 6)()
    err = C.create_string_buffer(600)
    rc = L.jt_thickness(h, v.ctypes.data, w, hh, d, int(inverse), int(mask), threshold, pixel_width,
                        out.ctypes.data if want_map else None, stats, info, err, 600, broken, broken_kind)
    return rc, out, list(stats), dict(zip(INFO, info)), err.value.decode()


def _both(L, h, vol, mask, want=(True, True), pixel_width=1.0):
    d, hh, w = vol.shape
    v = np.ascontiguousarray(vol, np.uint8)
    outs = [np.full(v.shape, -2.0, np.float32) if k else None for k in want]
    stats = (C.c_double * 10)()
    info = (C.c_int * 6)()
    err = C.create_string_buffer(600)
    rc = L.jt_thickness_both(h, v.ctypes.data, w, hh, d, int(mask), 128, pixel_width,
                             outs[0].ctypes.data if want[0] else None, outs[1].ctypes.data if want[1] else None,
                             stats, info, err, 600)
    return rc, outs, list(stats), dict(zip(INFO, info)), err.value.decode()


def test_create_without_a_device_throws():
    """no CUDA device: ThicknessNative.create throws RuntimeException carrying the library's message"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a CUDA device")
    from tests import jni_harness
    err = C.create_string_buffer(600)
    assert jni_harness.lib().jt_create(0, err, 600) == 0
    msg = err.value.decode()
    assert msg.startswith("java/lang/RuntimeException: ") and "no CUDA device" in msg and "no CPU fallback" in msg
    dev = (C.c_int * 2)(0, 1)
    assert jni_harness.lib().jt_create_multi(dev, 2, err, 600) == 0
    assert err.value.decode().startswith("java/lang/RuntimeException: ")


def _handle(emulated, multi=None):
    from tests import jni_harness
    L = jni_harness.lib(emulated=emulated)
    err = C.create_string_buffer(600)
    if multi is None:
        h = L.jt_create(0, err, 600)
    else:
        h = L.jt_create_multi((C.c_int * len(multi))(*multi), len(multi), err, 600)
    assert h != 0, err.value
    return L, h


@pytest.fixture(scope="module", params=[pytest.param("gpu", marks=pytest.mark.gpu), "emulated"])
def kind(request):
    return request.param


@pytest.fixture(scope="module")
def jni(kind):
    L, h = _handle(kind == "emulated")
    yield L, h
    L.jt_destroy(h)


@pytest.fixture(scope="module")
def jni_multi(kind):
    """ThicknessNative.createMulti over three slabs (the same device named three times: slabs may share a GPU)"""
    L, h = _handle(kind == "emulated", multi=[0, 0, 0])
    yield L, h
    L.jt_destroy(h)


def _check_map(out, st, ref):
    assert np.array_equal(np.isnan(out), np.isnan(ref["map"]))
    ok = ~np.isnan(out)
    assert np.allclose(out[ok], ref["map"][ok], rtol=1e-5, atol=0)          # float32 maps: 1e-5 relative (north star)
    assert int(st[0]) == ref["stats"].count
    assert st[1] == pytest.approx(ref["stats"].mean, rel=1e-5) and st[2] == pytest.approx(ref["stats"].sd, rel=1e-5)
    assert st[4] == pytest.approx(ref["stats"].max, rel=1e-5)


@pytest.mark.parametrize("inverse", [False, True])
def test_thickness_through_the_shim(jni, oracle, inverse):
    L, h = jni
    vol = shapes.random_blobs((13, 18, 21), 4)
    rc, out, st, info, err = _call(L, h, vol, inverse, True, pixel_width=0.5)
    assert rc == 1 and err == ""
    d = vol.shape[0]
    assert info["reads"] == d and info["writes"] == d and info["each_once"] == 1
    assert info["live_refs"] == 0 and info["max_live_refs"] <= 2 and info["violations"] == 0
    _check_map(out, st, oracle.process_image(vol, inverse=inverse, mask=True, pixel_width=0.5))


def test_statistics_only_call_reads_inputs_only(jni, oracle):
    L, h = jni
    vol = shapes.random_blobs((9, 12, 11), 2)
    rc, _, st, info, err = _call(L, h, vol, False, False, want_map=False)
    assert rc == 1 and err == ""
    assert info["reads"] == 9 and info["writes"] == 0 and info["live_refs"] == 0 and info["violations"] == 0
    assert int(st[0]) == oracle.process_image(vol, inverse=False, mask=False)["stats"].count


def test_not_binary_is_null_not_an_exception(jni):
    L, h = jni
    vol = np.zeros((3, 4, 5), np.uint8)
    vol[1, 2, 3] = 7
    rc, _, _, info, err = _call(L, h, vol, False, True)
    assert rc == 0 and err == ""                                             # processImage returns null
    assert info["live_refs"] == 0 and info["violations"] == 0 and info["writes"] == 0


def test_failure_is_a_runtime_exception(jni):
    L, h = jni
    vol = np.zeros((2, 3, 4), np.uint8)
    rc, _, _, info, err = _call(L, h, vol, False, True, dims=(2, 3, 0))     # w = 0
    assert rc == -1 and err.startswith("java/lang/RuntimeException: ") and "volume size" in err
    assert info["live_refs"] == 0 and info["violations"] == 0 and info["reads"] == 0


@pytest.mark.parametrize("how", [0, 1])
def test_null_or_short_slice_is_an_illegal_argument(jni, how):
    """validated before anything is copied: nothing was read, nothing written, no reference left"""
    L, h = jni
    vol = shapes.random_blobs((5, 6, 7), 1)
    rc, _, _, info, err = _call(L, h, vol, False, True, broken=3, broken_kind=how)
    assert rc == -1 and err.startswith("java/lang/IllegalArgumentException: ") and "slices[3]" in err
    assert info["reads"] == 0 and info["writes"] == 0 and info["live_refs"] == 0 and info["violations"] == 0


def test_dead_handle_is_an_illegal_argument(jni):
    L, _ = jni
    rc, _, _, info, err = _call(L, 0, np.zeros((2, 3, 4), np.uint8), False, True)
    assert rc == -1 and err.startswith("java/lang/IllegalArgumentException: ") and "handle" in err


@pytest.mark.parametrize("want", [(True, True), (False, True), (False, False)])
def test_both_through_the_shim(jni, oracle, want):
    """mapChoice = "Both" (ThicknessWrapper.java:127-134): one upload, {Tb.Th, Tb.Sp} statistics, maps on request"""
    L, h = jni
    vol = shapes.random_blobs((11, 16, 19), 7)
    rc, outs, st, info, err = _both(L, h, vol, True, want, pixel_width=0.7)
    assert rc == 1 and err == ""
    d = vol.shape[0]
    assert info["reads"] == d and info["writes"] == d * sum(want) and info["each_once"] == 1
    assert info["live_refs"] == 0 and info["violations"] == 0
    for k, inverse in enumerate((False, True)):
        ref = oracle.process_image(vol, inverse=inverse, mask=True, pixel_width=0.7)
        if want[k]:
            _check_map(outs[k], st[5 * k:5 * k + 5], ref)
        else:
            assert int(st[5 * k]) == ref["stats"].count and st[5 * k + 1] == pytest.approx(ref["stats"].mean, rel=1e-5)


def test_both_not_binary_is_null(jni):
    L, h = jni
    vol = np.full((3, 4, 5), 255, np.uint8)
    vol[0, 0, 0] = 1
    rc, _, _, info, err = _both(L, h, vol, True)
    assert rc == 0 and err == "" and info["writes"] == 0 and info["violations"] == 0


def test_create_multi_through_the_shim(jni, jni_multi, oracle):
    """the same calls on a createMulti handle: z-slabs inside the library, results those of one device bit for bit"""
    L1, h1 = jni
    L, h = jni_multi
    vol = shapes.random_blobs((23, 17, 20), 5)
    rc, outs, st, info, err = _both(L, h, vol, True, pixel_width=0.5)
    assert rc == 1 and err == "" and info["violations"] == 0 and info["live_refs"] == 0
    rc1, outs1, st1, _, _ = _both(L1, h1, vol, True, pixel_width=0.5)
    assert rc1 == 1
    for k in range(2):
        assert np.array_equal(outs[k].view(np.uint32), outs1[k].view(np.uint32))
        assert st[5 * k] == st1[5 * k] and st[5 * k + 3:5 * k + 5] == st1[5 * k + 3:5 * k + 5]
        assert st[5 * k + 1] == pytest.approx(st1[5 * k + 1], rel=1e-12) and st[5 * k + 2] == pytest.approx(st1[5 * k + 2], rel=1e-9)
    _check_map(outs[1], st[5:], oracle.process_image(vol, inverse=True, mask=True, pixel_width=0.5))
    rc, out, st, info, err = _call(L, h, vol, False, True)
    assert rc == 1 and info["violations"] == 0
    _check_map(out, st, oracle.process_image(vol, inverse=False, mask=True))


def test_create_multi_bad_arguments(jni):
    L, _ = jni
    err = C.create_string_buffer(600)
    assert L.jt_create_multi(None, -1, err, 600) == 0 and err.value.decode().startswith("java/lang/IllegalArgumentException: ")
    assert L.jt_create_multi((C.c_int * 1)(0), 0, err, 600) == 0 and err.value.decode().startswith("java/lang/IllegalArgumentException: ")
    assert L.jt_create_multi((C.c_int * 2)(0, 99), 2, err, 600) == 0 and err.value.decode().startswith("java/lang/RuntimeException: ")
