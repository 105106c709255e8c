"""The separable painter's per-line code (bonej2_b200/csrc/paint_sep_core.cuh), run line by line on the CPU
by oracle/sep_paint_model.cpp, against the literal Local_Thickness_Parallel restatement (oracle lto_paint).
The CUDA kernels execute the same statements; this keeps the algorithm itself under test without a GPU."""
import numpy as np
import pytest

from tests import shapes


def _ridge_u32(oracle, vol, inverse=False):
    dist, sq = oracle.edt(vol, inverse=inverse)
    rg = oracle.ridge(dist)
    return rg, np.where(rg > 0, sq, 0).astype(np.uint32)


def _check(oracle, ridge_f32, ridge_u32):
    _, want = oracle.paint(ridge_f32)
    got, info = oracle.sep_paint_model(ridge_u32)
    assert info["rc"] == 0 and info["overflow_act"] == 0 and info["overflow_front"] == 0
    assert np.array_equal(got, want)


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("inverse", [False, True])
def test_model_random_blobs(oracle, seed, inverse):
    vol = shapes.random_blobs((28, 33, 37), seed)
    _check(oracle, *_ridge_u32(oracle, vol, inverse))


@pytest.mark.parametrize("inverse", [False, True])
def test_model_phantom(oracle, inverse):
    from bonej2_b200.phantom import phantom_shapes
    vol = oracle.phantom_raster(phantom_shapes(56, 56, 56, 3), 56, 56, 56)
    _check(oracle, *_ridge_u32(oracle, vol, inverse))


def test_model_every_voxel_a_source(oracle):
    """no ridge pruning at all: every foreground voxel paints its ball (many equal tags, many ties)"""
    vol = shapes.random_blobs((20, 22, 24), 5)
    dist, sq = oracle.edt(vol)
    _check(oracle, dist, sq.astype(np.uint32))


def test_model_degenerate(oracle):
    z = np.zeros((5, 6, 7), np.uint32)
    got, _ = oracle.sep_paint_model(z)
    assert not got.any()
    one = z.copy(); one[2, 3, 4] = 9                       # a single ball of radius 3
    _check(oracle, np.sqrt(one.astype(np.float32)), one)
    line = np.zeros((1, 1, 40), np.uint32); line[0, 0, 5] = 16; line[0, 0, 30] = 100
    _check(oracle, np.sqrt(line.astype(np.float32)), line)


@pytest.mark.parametrize("na,nb", [(1, 1), (7, 3), (8, 4), (9, 5), (37, 28), (33, 37), (512, 128), (100, 1), (1, 100),
                                   (1024, 31), (250, 250)])
def test_patch_map_is_a_bijection(oracle, na, nb):
    """the kernels' thread -> line map (8 x 4 patches of lines per warp): every line exactly once, idle threads -1,
    and a warp's lines form one patch"""
    m = oracle.sep_patch_map(na, nb)
    assert m.size % 128 == 0
    used = m[m >= 0]
    assert used.size == na * nb and np.array_equal(np.sort(used), np.arange(na * nb))
    warps = m.reshape(-1, 32)
    for wl in warps[:: max(1, len(warps) // 64)]:
        v = wl[wl >= 0]
        if v.size == 0:
            continue
        a, b = v % na, v // na
        assert a.max() - a.min() < 8 and b.max() - b.min() < 4
        lanes = np.nonzero(wl >= 0)[0]
        assert np.array_equal(a - a.min() + (a.min() % 8), lanes % 8) and np.array_equal(b - (b.min() // 4) * 4, lanes // 8)
