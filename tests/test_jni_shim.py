"""The JNI half of the seam (bonej2_b200/csrc/jni_shim.cpp, bound by java/org/bonej/wrapperPlugins/b200/
ThicknessNative.java) driven by a fake JNIEnv, because the image has no JDK (tests/jni_harness.py).  Checks the shim's
contract from INTEGRATION.md: slices pinned and released exactly once (inputs discarded, outputs committed), the
input untouched, MapStats = the library's statistics, null for a non-binary image (LocalThicknessWrapper.processImage
returns null), RuntimeException with bjt_last_error() otherwise -- and never a CPU fallback."""
import ctypes as C

import numpy as np
import pytest

from tests import shapes


def _call(L, h, vol, inverse, mask, want_map=True, threshold=128, pixel_width=1.0, dims=None):
    d, hh, w = vol.shape if dims is None else dims
    v = np.ascontiguousarray(vol, np.uint8)
    out = np.full(v.shape, -2.0, np.float32) if want_map else None
    stats = (C.c_double * 5)()
    info = (C.c_int * 6)()
    err = C.create_string_buffer(600)
    rc = L.jt_thickness(h, v.ctypes.data, w, hh, d, int(inverse), int(mask), threshold, pixel_width,
                        out.ctypes.data if want_map else None, stats, info, err, 600)
    return rc, out, list(stats), dict(zip(("pins", "releases", "aborted", "committed", "still_pinned", "violations"), info)), err.value.decode()


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


@pytest.fixture(scope="module")
def jni():
    from tests import jni_harness
    L = jni_harness.lib()
    err = C.create_string_buffer(600)
    h = L.jt_create(0, err, 600)
    assert h != 0, err.value
    yield L, h
    L.jt_destroy(h)


@pytest.mark.gpu
@pytest.mark.parametrize("inverse", [False, True])
def test_thickness_through_the_shim(jni, oracle, inverse):
    L, h = jni
    vol = shapes.random_blobs((13, 18, 21), 4)
    rc, out, st, info, err = _call(L, h, vol, inverse, True, pixel_width=0.5)
    assert rc == 1 and err == ""
    d = vol.shape[0]
    assert info == dict(pins=2 * d, releases=2 * d, aborted=d, committed=d, still_pinned=0, violations=0)
    ref = oracle.process_image(vol, inverse=inverse, mask=True, pixel_width=0.5)
    assert np.array_equal(np.isnan(out), np.isnan(ref["map"]))
    ok = ~np.isnan(out)
    assert np.allclose(out[ok], ref["map"][ok], rtol=1e-5, atol=0)          # float32 maps: 1e-5 relative (north star)
    assert int(st[0]) == ref["stats"].count
    assert st[1] == pytest.approx(ref["stats"].mean, rel=1e-5) and st[2] == pytest.approx(ref["stats"].sd, rel=1e-5)
    assert st[4] == pytest.approx(ref["stats"].max, rel=1e-5)


@pytest.mark.gpu
def test_statistics_only_call_pins_inputs_only(jni, oracle):
    L, h = jni
    vol = shapes.random_blobs((9, 12, 11), 2)
    rc, _, st, info, err = _call(L, h, vol, False, False, want_map=False)
    assert rc == 1 and err == ""
    assert info == dict(pins=9, releases=9, aborted=9, committed=0, still_pinned=0, violations=0)
    assert int(st[0]) == oracle.process_image(vol, inverse=False, mask=False)["stats"].count


@pytest.mark.gpu
def test_not_binary_is_null_not_an_exception(jni):
    L, h = jni
    vol = np.zeros((3, 4, 5), np.uint8)
    vol[1, 2, 3] = 7
    rc, _, _, info, err = _call(L, h, vol, False, True)
    assert rc == 0 and err == ""                                             # processImage returns null
    assert info["still_pinned"] == 0 and info["violations"] == 0 and info["pins"] == info["releases"]


@pytest.mark.gpu
def test_failure_is_a_runtime_exception(jni):
    L, h = jni
    vol = np.zeros((2, 3, 4), np.uint8)
    rc, _, _, info, err = _call(L, h, vol, False, True, dims=(2, 3, 0))     # w = 0
    assert rc == -1 and err.startswith("java/lang/RuntimeException: ") and "volume size" in err
    assert info["still_pinned"] == 0 and info["violations"] == 0 and info["pins"] == info["releases"]
