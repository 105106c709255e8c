"""Hyperstack front-end: subspace order and labels as HyperstackUtilsTest.testViewMeta
(Modern/wrapperPlugins/src/test/java/org/bonej/wrapperPlugins/wrapperUtils/HyperstackUtilsTest.java:326-351),
and batched Thickness per subspace against the oracle (BASELINE.json configs[4])."""
import math

import numpy as np
import pytest

from bonej2_b200 import hyperstack as hs
from bonej2_b200.wrapper import CHANNEL, TIME, UNKNOWN, X, Y, Z
from tests import shapes


def test_view_meta_order_and_labels():
    data = np.zeros((2, 2, 2, 2, 2, 2), np.uint8)                 # numpy order: T, C, W, Z, Y, X
    axes = [X, Y, Z, UNKNOWN, CHANNEL, TIME]
    names = dict(hs.AXIS_NAMES)
    names[UNKNOWN] = "W"
    subs = list(hs.split_3d_subspaces(data, axes, names))
    assert [s.positions for s in subs] == [(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0), (0, 0, 1), (1, 0, 1),
                                           (0, 1, 1), (1, 1, 1)]
    assert [s.label for s in subs] == ["W: 0, Channel: 1, Time: 1", "W: 1, Channel: 1, Time: 1",
                                       "W: 0, Channel: 2, Time: 1", "W: 1, Channel: 2, Time: 1",
                                       "W: 0, Channel: 1, Time: 2", "W: 1, Channel: 1, Time: 2",
                                       "W: 0, Channel: 2, Time: 2", "W: 1, Channel: 2, Time: 2"]
    assert all(s.data.shape == (2, 2, 2) for s in subs)


def test_subspace_content_and_interleaved_axes():
    # ImageJ hyperstack order X, Y, C, Z, T
    c, zdim, t, h, w = 3, 4, 2, 5, 6
    data = np.arange(t * zdim * c * h * w, dtype=np.int64).reshape(t, zdim, c, h, w)
    subs = list(hs.split_3d_subspaces(data, [X, Y, CHANNEL, Z, TIME]))
    assert len(subs) == c * t
    assert subs[0].label == "Channel: 1, Time: 1" and subs[1].label == "Channel: 2, Time: 1"
    assert subs[3].label == "Channel: 1, Time: 2"
    assert np.array_equal(subs[4].data, data[1, :, 1])            # C = 2, T = 2 -> [z][y][x]


def test_same_type_axes_get_subscripts():
    data = np.zeros((2, 2, 3, 3, 3), np.uint8)
    subs = list(hs.split_3d_subspaces(data, [X, Y, Z, CHANNEL, CHANNEL]))
    assert subs[1].label == "Channel: 2, Channel(2): 1"


@pytest.mark.gpu
def test_batched_thickness_matches_oracle(oracle):
    vols = [shapes.random_blobs((12, 14, 16), s) for s in (1, 2, 3)] + [np.zeros((12, 14, 16), np.uint8)]
    data = np.stack(vols).reshape(2, 2, 12, 14, 16)               # T, C, Z, Y, X
    rows = hs.thickness_hyperstack(data, [X, Y, Z, CHANNEL, TIME], name="stack", map_choice="Both", unit="mm")
    assert [r[0] for r in rows] == ["stack Channel: 1, Time: 1", "stack Channel: 2, Time: 1",
                                    "stack Channel: 1, Time: 2", "stack Channel: 2, Time: 2"]
    for (label, vals), vol in zip(rows, vols):
        for prefix, inverse in (("Tb.Th", False), ("Tb.Sp", True)):
            st = oracle.process_image(vol, inverse=inverse, mask=True)["stats"]
            got = vals[f"{prefix} Mean (mm)"]
            if st.count == 0:
                assert math.isnan(got)
            else:
                assert math.isclose(got, st.mean, rel_tol=1e-5)
                assert math.isclose(vals[f"{prefix} Max (mm)"], st.max, rel_tol=1e-5)
