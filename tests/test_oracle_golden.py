"""Pins the CPU oracle (oracle/lt_oracle.c) to the reference's own known-answer tests:
  ThicknessWrapperTest.java:225-259  2x2x2 zeros, Both, mask off -> NaN x3, 10.392304420471191, 0, 10.39...
  ThicknessHelperTest.java:45-78     rods +-1.5, spheres +-10 %, bricks +-5 %
and to the probe-derived vectors of SURVEY.md Appendix B (regression anchors)."""
import math

import numpy as np
import pytest

from tests import shapes


def test_wrapper_kat_2x2x2(oracle):
    v = np.zeros((2, 2, 2), np.uint8)
    th = oracle.process_image(v, inverse=False, mask=False)
    sp = oracle.process_image(v, inverse=True, mask=False)
    assert th["stats"].count == 0                      # -> NaN, NaN, NaN in ThicknessWrapper.java:170-175
    assert np.isnan(th["map"]).all()
    assert sp["stats"].mean == 10.392304420471191
    assert sp["stats"].sd == 0.0
    assert sp["stats"].max == 10.392304420471191
    assert sp["stats"].count == 8


@pytest.mark.parametrize("d", list(range(1, 25)))          # the reference's full range (length 20 d here, 100 d through the C-ABI: tests/test_gpu_kat.py)
def test_rods(oracle, d):
    r = oracle.process_image(shapes.rod(d, 20 * d), inverse=False, mask=False)
    m = r["map"]
    assert abs(np.nanmean(m[m > 0]) - d) <= 1.5


@pytest.mark.parametrize("r", list(range(2, 25)))          # ThicknessHelperTest.java:56-65, full range
def test_spheres(oracle, r):
    res = oracle.process_image(shapes.sphere(r), inverse=False, mask=False)
    m = res["map"]
    expected = 1.9441872882 * r - 1.218936
    assert abs(np.nanmean(m[m > 0]) - expected) <= 0.1 * expected


ANCHORS = {1: 2.0, 2: 2.0, 3: 4.0, 4: 4.0, 5: 5.9410, 8: 7.8921, 11: 11.7252, 20: 19.2932}


@pytest.mark.parametrize("t", list(range(1, 21)))           # ThicknessHelperTest.java:67-78, full range
def test_bricks(oracle, t):
    expected = ANCHORS.get(t)
    res = oracle.process_image(shapes.brick(t, 128), inverse=False, mask=False)
    m = res["map"]
    mean = np.nanmean(m[m > 0])
    target = t if t % 2 == 0 else t + 1
    assert abs(mean - target) <= 0.05 * target + (1 if t <= 2 else 0)   # reference tolerance (t=1,2 -> 2)
    if expected is not None:
        assert abs(mean - expected) < 2e-4                               # Appendix B anchors


@pytest.mark.parametrize("inverse,mask,mean,sd,mx,n", [
    (False, True, 9.34699221, 1.25715655, 10.0, 1000),
    (True, True, 2.44035259, 0.492762451, 3.46410155, 728),
    (False, False, 8.02187012, 2.18764581, 10.0, 1600),
    (True, False, 2.35687182, 0.423835819, 3.46410155, 1216),
])
def test_hyper_cube_anchors(oracle, inverse, mask, mean, sd, mx, n):
    st = oracle.process_image(shapes.hyper_cube(), inverse=inverse, mask=mask)["stats"]
    assert st.count == n
    assert math.isclose(st.mean, mean, rel_tol=1e-7)
    assert math.isclose(st.sd, sd, rel_tol=1e-7)
    assert math.isclose(st.max, mx, rel_tol=1e-7)


def test_not_binary_returns_none(oracle):
    v = np.zeros((3, 3, 3), np.uint8)
    v[1, 1, 1] = 7
    assert oracle.process_image(v) is None


def test_paint_invariant_small(oracle):
    """A.3 invariant: painting from the ridge equals painting from every foreground voxel."""
    vol = shapes.random_blobs((20, 22, 24), 5)
    dist, sq = oracle.edt(vol)
    rg = oracle.ridge(dist)
    _, from_ridge = oracle.paint(rg)
    _, from_all = oracle.paint(dist)
    assert np.array_equal(from_ridge, from_all)
    # direct evaluation of LT^2(p) = max{ sq(q) : |p-q|^2 <= sq(q) }
    zz, yy, xx = np.nonzero(sq)
    out = np.zeros(sq.shape, np.uint32)
    Z, Y, X = np.indices(sq.shape)
    for z, y, x in zip(zz, yy, xx):
        R = sq[z, y, x]
        m = ((Z - z) ** 2 + (Y - y) ** 2 + (X - x) ** 2) <= R
        out[m] = np.maximum(out[m], R)
    assert np.array_equal(out, from_ridge)


@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("shape,seed", [((28, 33, 37), 1), ((40, 24, 31), 2), ((17, 50, 23), 3)])
def test_edt_against_scipy(oracle, shape, seed, inverse):
    """Independent pin of the EDT restatement: SciPy's exact Euclidean distance transform (a different
    algorithm and code base) gives the same squared distances wherever the volume has background; the
    sentinel 3 (n + 1)^2 of EDT_S1D only appears when there is none."""
    ndi = pytest.importorskip("scipy.ndimage")
    vol = shapes.random_blobs(shape, seed)
    _, sq = oracle.edt(vol, inverse=inverse)
    fg = (vol < 128) if inverse else (vol >= 128)
    assert (~fg).any()
    want = np.rint(ndi.distance_transform_edt(fg) ** 2).astype(np.int64)
    assert np.array_equal(sq.astype(np.int64), want)


def test_edt_without_background_is_the_sentinel(oracle):
    vol = np.full((5, 6, 7), 255, np.uint8)
    _, sq = oracle.edt(vol)
    assert (sq == 3 * (7 + 1) ** 2).all()          # n = the largest extent


@pytest.mark.parametrize("shape,density", [((20, 30, 40), 0.5), ((64, 50, 33), 0.02), ((33, 47, 64), 0.999), ((1, 1, 50), 0.3),
                                           ((7, 1, 1), 0.5), ((12, 9, 5), 1.0), ((12, 9, 5), 0.0)])
@pytest.mark.parametrize("inverse", [False, True])
def test_exact_edt_checker_matches_the_restatement(oracle, shape, density, inverse):
    """oracle/edt_exact.c (lower envelope of parabolas, O(N)) against lto_edt (the reference's brute-force scans): the
    large-volume checker of tests/test_gpu_configs.py is pinned here, where both can run"""
    rng = np.random.default_rng(17)
    vol = np.where(rng.random(shape) < density, 255, 0).astype(np.uint8)
    _, sq = oracle.edt(vol, inverse=inverse)
    assert np.array_equal(oracle.edt_exact_sq(vol, inverse=inverse), sq)
    assert np.array_equal(oracle.edt_exact_sq(vol, inverse=inverse, nthreads=1), sq)


def test_exact_edt_checker_on_the_phantom(oracle):
    from bonej2_b200.phantom import phantom_shapes
    vol = oracle.phantom_raster(phantom_shapes(96, 96, 96, 1), 96, 96, 96)
    for inverse in (False, True):
        assert np.array_equal(oracle.edt_exact_sq(vol, inverse=inverse), oracle.edt(vol, inverse=inverse)[1])
