"""Test volumes: the known-answer shapes of the reference's own tests plus edge cases.
Rod / sphere / brick construction follows Legacy/bonej/src/test/java/org/bonej/plugins/
ThicknessHelperTest.java:88-194 (ImageJ fillOval is approximated by an exact disc test, the
tolerances of the reference tests absorb the difference)."""
import numpy as np


def _disc(side_y, side_x, cy, cx, diam):
    """ImageJ-like filled oval of bounding box [cx, cx+diam) x [cy, cy+diam)"""
    yy, xx = np.mgrid[0:side_y, 0:side_x]
    r = diam / 2.0
    return ((xx + 0.5 - (cx + r)) ** 2 + (yy + 0.5 - (cy + r)) ** 2) <= r * r


def rod(d, length=None):
    length = length or 10 * d
    side = 2 * d
    sl = _disc(side, side, d // 2, d // 2, d)
    vol = np.zeros((length, side, side), np.uint8)
    vol[:, sl] = 255
    return vol


def sphere(r):
    side = 2 * r + 2
    vol = np.zeros((side, side, side), np.uint8)
    for z in range(-r, r + 1):
        rc = int(round(np.sqrt(r * r - z * z)))
        if rc > 0:
            vol[z + r + 1][_disc(side, side, r + 1 - rc, r + 1 - rc, 2 * rc)] = 255
    return vol


def brick(t, wh=32):
    vol = np.zeros((t + 2, wh + 2, wh + 2), np.uint8)
    vol[1:t + 1, 1:wh + 1, 1:wh + 1] = 255
    return vol


def hyper_cube():
    """one subspace of IntegrationTestLogs hyperCube.tif: a 10^3 cube inside 12^3"""
    vol = np.zeros((12, 12, 12), np.uint8)
    vol[1:11, 1:11, 1:11] = 255
    return vol


def random_blobs(shape, seed, p=0.35, smooth=2):
    rng = np.random.default_rng(seed)
    a = rng.random(shape)
    for ax in range(3):
        for _ in range(smooth):
            a = (a + np.roll(a, 1, ax) + np.roll(a, -1, ax)) / 3.0
    thr = np.quantile(a, 1.0 - p)
    return np.where(a > thr, 255, 0).astype(np.uint8)


def edge_cases():
    """name -> volume [z][y][x]"""
    cases = {}
    cases["2x2x2_zero"] = np.zeros((2, 2, 2), np.uint8)
    cases["2x2x2_full"] = np.full((2, 2, 2), 255, np.uint8)
    cases["1x1x1_fg"] = np.full((1, 1, 1), 255, np.uint8)
    one_bg = np.full((9, 11, 13), 255, np.uint8)
    one_bg[4, 5, 6] = 0
    cases["single_bg_voxel"] = one_bg
    one_fg = np.zeros((7, 8, 9), np.uint8)
    one_fg[3, 4, 5] = 255
    cases["single_fg_voxel"] = one_fg
    cases["noncubic_37x101x13"] = random_blobs((13, 101, 37), 7)
    cases["slice_2d_d1"] = random_blobs((1, 64, 48), 3)          # the per-slice 2-D case of SliceGeometry
    cases["row_1d"] = random_blobs((1, 1, 200), 5, smooth=4)
    cases["hyper_cube"] = hyper_cube()
    cases["brick_t5"] = brick(5)
    cases["sphere_r6"] = sphere(6)
    cases["rod_d7"] = rod(7, 40)
    cases["blobs_40"] = random_blobs((40, 40, 40), 11)
    cases["wide_1100"] = random_blobs((3, 5, 1100), 13, smooth=6)   # two 1024-voxel row segments
    planes = np.zeros((30, 20, 20), np.uint8)
    planes[0:9:2] = 255
    cases["aniso_planes"] = planes                                # AnisotropicHyperStack-like plates
    return cases
