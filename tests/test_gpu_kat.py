"""The reference's own known-answer ranges through the C-ABI, with the reference's own tolerances
(Legacy/bonej/src/test/java/org/bonej/plugins/ThicknessHelperTest.java:45-78): rods d = 1..24 of length 100 d
(mean = d +- 1.5), spheres r = 2..24 (mean = 1.9441872882 r - 1.218936 +- 10 %), bricks t = 1..20 of 128 x 128 (mean =
t, or t + 1 for odd t, +- 5 %).  ThicknessHelper.getLocalThickness(imp, false) = foreground, no mask, calibrated; the
statistic is ThicknessHelperTest.meanStdDev (:118-156): mean over the pixels > 0.  The shapes are built as the test
builds them (tests/shapes.py; ImageJ's fillOval is approximated by an exact disc, which the tolerances absorb).
Every squared EDT is in addition compared bit for bit with the independent exact checker (oracle/edt_exact.c)."""
import numpy as np
import pytest

from tests import shapes

pytestmark = pytest.mark.gpu


def _mean_positive(m):
    v = m[np.nan_to_num(m, nan=0.0) > 0]
    return float(v.astype(np.float64).mean())


def _run(ctx, oracle, vol):
    m, st, tm = ctx.thickness(vol, inverse=False, mask=False)
    assert np.array_equal(ctx.edt_sq(vol, inverse=False), oracle.edt_exact_sq(vol, inverse=False))
    mean = _mean_positive(m)
    assert st.mean == pytest.approx(mean, rel=1e-6)          # StackStatistics over the non-NaN voxels: the same set
    return mean


@pytest.mark.parametrize("d", list(range(1, 25)))
def test_rods_full_range(ctx, oracle, d):
    assert abs(_run(ctx, oracle, shapes.rod(d, 100 * d)) - d) <= 1.5


@pytest.mark.parametrize("r", list(range(2, 25)))
def test_spheres_full_range(ctx, oracle, r):
    regression = 1.9441872882 * r - 1.218936
    assert abs(_run(ctx, oracle, shapes.sphere(r)) - regression) <= 0.1 * regression


@pytest.mark.parametrize("t", list(range(1, 21)))
def test_bricks_full_range(ctx, oracle, t):
    expected = t + 1 if t % 2 else t
    assert abs(_run(ctx, oracle, shapes.brick(t, 128)) - expected) <= 0.05 * expected


@pytest.mark.parametrize("r", [2, 7, 16, 24])
def test_spheres_equal_the_oracle(ctx, oracle, r):
    """and the map itself against the restatement, value for value"""
    vol = shapes.sphere(r)
    m, st, _ = ctx.thickness(vol, inverse=False, mask=False)
    o = oracle.process_image(vol, inverse=False, mask=False)
    assert np.array_equal(np.isnan(m), np.isnan(o["map"]))
    ok = ~np.isnan(m)
    assert np.allclose(m[ok], o["map"][ok], rtol=1e-5, atol=0)   # float32 maps: 1e-5 relative (north star)
    assert st.count == o["stats"].count
