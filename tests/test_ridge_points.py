"""FindRidgePoints: the reference's known-answer test for the op on both the CPU restatement
and the CUDA path, and the CUDA path against the restatement on random volumes."""
import numpy as np
import pytest

from oracle import find_ridge_oracle as fro
from tests import shapes


def _sphere_image():
    """FindRidgePointsTest.getSphereImage (FindRidgePointsTest.java:70-92): radius 10 about (50, 50, 50) in 101^3"""
    z, y, x = np.indices((101, 101, 101))
    return (((x - 50) ** 2 + (y - 50) ** 2 + (z - 50) ** 2) <= 100).astype(np.uint8) * 255


def test_oracle_sphere_ridge():                                      # FindRidgePointsTest.java:49-66
    pts, _ = fro.find_ridge_points(_sphere_image())
    assert pts.shape == (1, 3) and pts[0].tolist() == [50.5, 50.5, 50.5]


@pytest.mark.gpu
def test_cuda_sphere_ridge(ctx):
    pts, n, mx = ctx.find_ridge_points(_sphere_image())
    assert n == 1 and pts[0].tolist() == [50.5, 50.5, 50.5] and mx > 0


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("fraction", [0.6, 0.2])
def test_cuda_matches_restatement(ctx, seed, fraction):
    vol = shapes.random_blobs((30, 34, 38), seed)
    want, wmx = fro.find_ridge_points(vol, fraction)
    got, n, mx = ctx.find_ridge_points(vol, fraction)
    assert mx == pytest.approx(wmx, rel=1e-6)
    assert n == len(want)
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_cuda_degenerate(ctx):
    pts, n, mx = ctx.find_ridge_points(np.zeros((5, 6, 7), np.uint8))
    assert n == 0 and mx == 0 and pts.shape == (0, 3)
    full = np.full((9, 9, 9), 255, np.uint8)                          # the zero border makes it a cube of distance values
    want, _ = fro.find_ridge_points(full)
    got, n, _ = ctx.find_ridge_points(full)
    assert n == len(want) and np.array_equal(got, want)
    pts, n, _ = ctx.find_ridge_points(full, cap=1)                    # more found than fit: count still reported
    assert n == len(want) and pts.shape == (min(1, n), 3)
