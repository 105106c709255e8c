"""Segmented statistics of a thickness map against direct restatements of the Java loops:
ParticleAnalysis.getMeanStdDev (Legacy/bonej/src/main/java/org/bonej/plugins/ParticleAnalysis.java:325-364) and
the per-slice cortical-thickness loop of SliceGeometry (.../SliceGeometry.java:876-900, 1172-1200).
Tolerance 1e-12 relative: sums are formed in double in a different order."""
import numpy as np
import pytest

from tests import shapes

pytestmark = pytest.mark.gpu


def _mean_std_dev(values, labels, sizes):
    """the two passes of getMeanStdDev, vectorised, same arithmetic (double sums of the float values > 0)"""
    n = len(sizes)
    v = values.astype(np.float64).reshape(-1)
    lab = labels.reshape(-1)
    pos = v > 0
    out = np.zeros((n, 3))
    sums = np.bincount(lab[pos], weights=v[pos], minlength=n)
    with np.errstate(divide="ignore", invalid="ignore"):
        out[1:, 0] = sums[1:] / sizes[1:]
        res = v[pos] - out[lab[pos], 0]
        ssq = np.bincount(lab[pos], weights=res * res, minlength=n)
        np.maximum.at(out[:, 2], lab[pos], v[pos])
        out[1:, 1] = np.sqrt(ssq[1:] / sizes[1:])
    return out


def _labels_for(vol, seed):
    """a label image: 0 = background, foreground voxels carved into a few 'particles' by position"""
    rng = np.random.default_rng(seed)
    d, h, w = vol.shape
    z, y, x = np.indices(vol.shape)
    lab = 1 + (x * 3 // w) + 3 * (y * 2 // h) + 6 * (z * 2 // d)
    lab = np.where(vol > 0, lab, 0).astype(np.int32)
    lab[rng.random(vol.shape) < 0.01] = 0                       # some foreground voxels labelled background
    return lab, int(lab.max()) + 2                                # one label that never occurs (size 0)


@pytest.mark.parametrize("seed", [1, 2])
def test_label_stats_match_get_mean_std_dev(ctx, seed):
    vol = shapes.random_blobs((40, 36, 44), seed)
    m, _, _ = ctx.thickness(vol, inverse=False, mask=True, pixel_width=0.5)      # NaN outside the foreground
    lab, n = _labels_for(vol, seed)
    sizes = np.bincount(lab.reshape(-1), minlength=n).astype(np.int64)
    got = ctx.label_stats(m, lab, sizes)
    want = _mean_std_dev(m, lab, sizes)
    assert got.shape == (n, 3)
    assert np.allclose(got[:-1], want[:-1], rtol=1e-12, atol=0)
    assert np.isnan(got[-1, 0]) and np.isnan(want[-1, 0]) and got[-1, 2] == 0      # empty particle: 0 / 0 as in Java
    assert got[0, 0] == 0 and got[0, 1] == 0                                      # label 0 is never filled


def test_label_stats_rejects_unknown_label(ctx):
    m = np.ones((2, 2, 2), np.float32)
    lab = np.full((2, 2, 2), 5, np.int32)
    with pytest.raises(RuntimeError, match="label outside"):
        ctx.label_stats(m, lab, np.array([0, 8], np.int64))


@pytest.mark.parametrize("roi", [None, (5, 3, 20, 17), (0, 0, 1, 1)])
def test_slice_stats_match_slice_geometry(ctx, roi):
    vol = shapes.random_blobs((24, 30, 36), 7)
    vol[5] = 0                                                   # an empty slice
    m, _, _ = ctx.thickness(vol, inverse=False, mask=True, pixel_width=1.0)
    got, cnt = ctx.slice_stats(m, roi)
    d, h, w = vol.shape
    rx, ry, rw, rh = roi if roi else (0, 0, w, h)
    for z in range(d):
        px = m[z, ry:ry + rh, rx:rx + rw].astype(np.float64).reshape(-1)
        px = px[px > 0]
        assert cnt[z] == px.size
        if px.size == 0:
            assert np.isnan(got[z, 0]) and got[z, 1] == 0 and np.isnan(got[z, 2])
            continue
        mean = px.sum() / px.size
        assert got[z, 0] == pytest.approx(mean, rel=1e-12)
        assert got[z, 1] == px.max()
        assert got[z, 2] == pytest.approx(np.sqrt(((mean - px) ** 2).sum() / px.size), rel=1e-9, abs=1e-12)


def test_slice_stats_bad_roi(ctx):
    m = np.ones((2, 4, 4), np.float32)
    with pytest.raises(RuntimeError, match="ROI"):
        ctx.slice_stats(m, (2, 2, 4, 1))
