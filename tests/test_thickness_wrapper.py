"""Mirror of Modern/wrapperPlugins/src/test/java/org/bonej/wrapperPlugins/ThicknessWrapperTest.java
(:72-281) and the shared cancel checks of CommonWrapperTests.java:81-224, driven through the C++
ThicknessWrapper (include/thickness_wrapper.hpp).  Validation runs before any device work, so the
cancel tests need no GPU; the tests that produce maps are marked gpu."""
import math

import numpy as np
import pytest

from bonej2_b200 import wrapper as W


@pytest.fixture(autouse=True)
def _reset_table():
    W._native.build()
    W.shared_table_reset()        # AbstractWrapperTest.java: SharedTable.reset() after each test
    yield
    W.shared_table_reset()


def test_null_image_cancels():
    o = W.run_thickness(None)
    assert o.canceled and o.cancel_reason == "No image open"


def test_2d_image_cancels():
    o = W.run_thickness(W.Dataset(np.zeros((10, 10), np.uint8), axes=[W.X, W.Y]))
    assert o.canceled and o.cancel_reason == "Need a 3D (X, Y, Z) image"


def test_composite_image_cancels():
    # ThicknessWrapperTest.testCompositeImageCancelsPlugin: 3x3x3 with 3 channels
    ds = W.Dataset(np.zeros((3, 3, 3, 3), np.uint8), axes=[W.X, W.Y, W.CHANNEL, W.Z])
    o = W.run_thickness(ds)
    assert o.canceled and o.cancel_reason == "Image cannot be composite. Please split the channels."


def test_time_dimension_cancels():
    ds = W.Dataset(np.zeros((3, 3, 3, 3), np.uint8), axes=[W.X, W.Y, W.Z, W.TIME])
    o = W.run_thickness(ds)
    assert o.canceled and o.cancel_reason == "Image cannot have time axis. Please split the hyperstack."


def test_non_binary_cancels():
    v = np.arange(125, dtype=np.uint8).reshape(5, 5, 5)       # the 5x5x5 ramp of CommonWrapperTests
    o = W.run_thickness(W.Dataset(v))
    assert o.canceled and o.cancel_reason == "Need a binary image"


def test_non_8bit_cancels():
    o = W.run_thickness(W.Dataset(np.zeros((4, 4, 4), np.uint16)))
    assert o.canceled and o.cancel_reason == "Need a binary image"


def test_anisotropy_dialog_cancel():
    # CommonWrapperTests.testAnisotropyWarning: pixelWidth 300 on a 5^3 image, user answers CANCEL
    ds = W.Dataset(np.zeros((5, 5, 5), np.uint8), scales=[300.0, 1.0, 1.0], units=["mm"] * 3)
    o = W.run_thickness(ds, headless=False, dialog_ok=False)
    assert o.canceled and o.cancel_reason == ""


@pytest.mark.gpu
def test_anisotropy_headless_warns_and_continues():
    ds = W.Dataset(np.zeros((5, 5, 5), np.uint8), scales=[300.0, 1.0, 1.0], units=["mm"] * 3)
    o = W.run_thickness(ds, headless=True, mapChoice="Trabecular separation")
    assert not o.canceled
    assert "The image is anisotropic (29900.0 %)" in o.log


@pytest.mark.gpu
def test_bad_map_choice_throws():
    o = W.run_thickness(W.Dataset(np.zeros((2, 2, 2), np.uint8)), mapChoice="Cortical thickness")
    assert o.exception == 1 and "Unexpected map choice" in o.cancel_reason


@pytest.mark.gpu
def test_results_kat():
    """ThicknessWrapperTest.testResults (:225-259)"""
    ds = W.Dataset(np.zeros((2, 2, 2), np.uint8), units=["mm"] * 3, name="kat")
    o = W.run_thickness(ds, mapChoice="Both", maskArtefacts=False, showMaps=False)
    assert o.headers == ["Tb.Th Mean (mm)", "Tb.Th Std Dev (mm)", "Tb.Th Max (mm)", "Tb.Sp Mean (mm)",
                         "Tb.Sp Std Dev (mm)", "Tb.Sp Max (mm)"]
    assert o.row_labels == ["kat"]
    exp = [math.nan, math.nan, math.nan, 10.392304420471191, 0.0, 10.392304420471191]
    for h, e in zip(o.headers, exp):
        v = o.table[h][0]
        assert (math.isnan(v) and math.isnan(e)) or v == e, (h, v, e)
    assert o.trabecularMap is None and o.separationMap is None       # showMaps = false -> null outputs


@pytest.mark.gpu
@pytest.mark.parametrize("choice,th,sp", [("Trabecular thickness", True, False), ("Trabecular separation", False, True),
                                           ("Both", True, True)])
def test_map_outputs(choice, th, sp):
    """ThicknessWrapperTest :130-210: output presence per mapChoice, FIRE LUT, input not overwritten"""
    vol = np.zeros((2, 2, 2), np.uint8)
    keep = vol.copy()
    o = W.run_thickness(W.Dataset(vol), mapChoice=choice, showMaps=True)
    assert (o.trabecularMap is not None) == th and (o.separationMap is not None) == sp
    assert o.color_table == "FIRE"
    assert np.array_equal(vol, keep)
    if sp:
        assert o.separationMap.dtype == np.float32 and o.separationMap.shape == vol.shape
        assert o.separation_max == pytest.approx(10.392304420471191)


@pytest.mark.gpu
def test_default_unit_and_rows_accumulate():
    vol = np.zeros((3, 3, 3), np.uint8)
    vol[1, 1, 1] = 255
    o1 = W.run_thickness(W.Dataset(vol, name="a"), mapChoice="Trabecular thickness", showMaps=False)
    assert o1.headers == ["Tb.Th Mean (pixel)", "Tb.Th Std Dev (pixel)", "Tb.Th Max (pixel)"]
    o2 = W.run_thickness(W.Dataset(vol, name="a"), mapChoice="Trabecular thickness", showMaps=False)
    assert o2.row_labels == ["a", "a"]                      # same label, cells taken -> a new row (SharedTable policy)
    o3 = W.run_thickness(W.Dataset(vol, name="a"), mapChoice="Trabecular separation", showMaps=False)
    assert o3.row_labels == ["a", "a"] and len(o3.headers) == 6   # fills the empty Tb.Sp cells of the last "a" row
    assert not math.isnan(o3.table["Tb.Sp Mean (pixel)"][1]) and math.isnan(o3.table["Tb.Sp Mean (pixel)"][0])
