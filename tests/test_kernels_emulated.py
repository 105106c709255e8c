"""The CUDA kernels' own source, executed on the CPU.  tests/cuda_emu/ compiles bonej2_b200/csrc/{edt,ridge,paint_sep}.cu
with g++ against a small emulation of the CUDA constructs they use (blocks one after another, the threads of a block
as fibers, barriers and warp collectives as rendezvous points; every array the kernels see ends or starts at an
inaccessible page).  What it checks is kernel LOGIC against the oracle without a GPU -- staging loops, tile and
patch index arithmetic, the warp-aggregated list allocation, redo paths, column blocks -- at sizes the emulation
finishes in seconds.  It complements, and does not replace, the `-m gpu` parity tests."""
import numpy as np
import pytest

from tests import kernels_emu as emu
from tests import shapes


@pytest.fixture(scope="module", params=[0, 1], ids=["guard_behind", "guard_in_front"])
def guard(request):
    emu.set_guard_mode(request.param)
    yield request.param
    emu.set_guard_mode(0)


@pytest.fixture(scope="module")
def guard_behind():
    """the painter's lists live inside one workspace allocation, so a second guard position buys little there"""
    emu.set_guard_mode(0)
    yield 0


def _ridge_of(oracle, vol, inverse):
    dist, sq = oracle.edt(vol, inverse=inverse)
    rg = oracle.ridge(dist)
    return rg, np.where(rg > 0, sq, 0).astype(np.uint32), sq


# (d, h, w): w = 32 / 64 -> tile kernel (32 columns per block); h or d = 700 -> 16 columns per block (shared-memory
# budget); w = 12 -> four-column line kernel; w = 13 -> one-thread-per-line kernel; 1100 -> two row segments in the x pass
EDT_SHAPES = [(9, 20, 32), (5, 33, 64), (1, 700, 32), (700, 1, 32), (7, 9, 12), (6, 5, 13), (3, 4, 1100), (1, 1, 40), (2, 2, 2)]


@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("shape", EDT_SHAPES)
def test_edt_kernels(oracle, guard, shape, inverse):
    vol = shapes.random_blobs(shape, 7)
    _, want = oracle.edt(vol, inverse=inverse)
    got, mx = emu.edt_sq(vol, inverse=inverse)
    assert np.array_equal(got, want)
    assert mx == int(want.max())


@pytest.mark.parametrize("shape", [(1, 100, 32), (100, 1, 32), (1, 1, 300)])
def test_edt_kernels_far_from_the_background(oracle, guard, shape):
    """a single background voxel: the y / z walks run far beyond the 64 sentinel rows either side of the staged lines"""
    vol = np.full(shape, 255, np.uint8)
    vol[0, 0, 3] = 0
    _, want = oracle.edt(vol)
    got, mx = emu.edt_sq(vol)
    assert np.array_equal(got, want) and mx == int(want.max())


@pytest.mark.parametrize("shape", [(1, 1500, 32), (3, 40, 48), (2, 37, 96)])
def test_edt_pair_kernel_tile_widths(oracle, guard_behind, shape):
    """the 16-bit pair kernel (edt_tile16_kernel): 16 columns per tile where the lines are too long for 32 (1500) or the
    width is a multiple of 16 only (48), groups of three positions that run past the end of a line (37, 40)"""
    for inverse in (False, True):
        vol = shapes.random_blobs(shape, 11)
        _, want = oracle.edt(vol, inverse=inverse)
        got, mx = emu.edt_sq(vol, inverse=inverse)
        assert np.array_equal(got, want) and mx == int(want.max())


def test_edt_pair_kernel_redo_of_a_16_column_tile(oracle, guard_behind):
    """a width of 16 with one background voxel: a 16-column pair tile (what lines of 2048 take as well) whose walks
    leave the pad rows, redone 8 columns at a time in 32 bits"""
    vol = np.full((1, 150, 16), 255, np.uint8)
    vol[0, 7, 5] = 0
    _, want = oracle.edt(vol)
    got, mx = emu.edt_sq(vol)
    assert np.array_equal(got, want) and mx == int(want.max())


def test_edt_pair_kernel_redoes_only_what_saturates(oracle, guard_behind):
    """tiles of short walks beside tiles whose walks leave the pad rows (the block redoes the tile in 32 bits), values
    above the 16-bit clamp (55000) included; and a tile in which only a few columns saturate"""
    vol = np.full((1, 16, 288), 255, np.uint8)
    vol[0, ::3, :2] = 0                                                  # background in the first two columns only
    _, want = oracle.edt(vol)                                            # x distance up to 286: y walks beyond column ~96 run through the pad
    got, mx = emu.edt_sq(vol)
    assert want.max() > 55000 and np.array_equal(got, want) and mx == int(want.max())
    vol = np.full((12, 1, 128), 255, np.uint8)                           # the same along z
    vol[::2, 0, :1] = 0
    _, want = oracle.edt(vol)
    got, mx = emu.edt_sq(vol)
    assert want.max() > 96 * 96 and np.array_equal(got, want) and mx == int(want.max())


def test_edt_kernels_no_background_and_no_foreground(oracle, guard):
    full = np.full((4, 6, 32), 255, np.uint8)
    got, mx = emu.edt_sq(full)
    assert (got == 3 * 33 * 33).all() and mx == 3 * 33 * 33
    got, mx = emu.edt_sq(np.zeros((4, 6, 32), np.uint8))
    assert not got.any() and mx == 0


@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("shape", [(9, 20, 32), (11, 13, 35), (8, 8, 31), (1, 30, 40), (17, 3, 5)])
def test_ridge_kernel(oracle, guard, shape, inverse):
    vol = shapes.random_blobs(shape, 5)
    _, want, sq = _ridge_of(oracle, vol, inverse)
    if not sq.any():
        pytest.skip("no foreground")
    got, count = emu.ridge(sq.astype(np.uint32))
    assert np.array_equal(got, want)
    assert count == int((want > 0).sum())


# partial 8 x 4 patches of lines in every stage (extents that are not multiples of 8 / 4), one line, one plane
PAINT_SHAPES = [(9, 20, 32), (13, 11, 7), (5, 37, 33), (6, 6, 6), (1, 9, 17), (10, 1, 12), (3, 5, 1)]


@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("shape", PAINT_SHAPES)
def test_painter_kernels(oracle, guard_behind, shape, inverse):
    vol = shapes.random_blobs(shape, 3)
    rg, ridge, _ = _ridge_of(oracle, vol, inverse)
    _, want = oracle.paint(rg)
    got, flags, _ = emu.paint(ridge)
    assert flags == 0 and np.array_equal(got, want)


@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("shape", PAINT_SHAPES + [(24, 40, 70), (40, 7, 33)])
def test_disc_painter_kernels(oracle, guard_behind, shape, inverse):
    """paint path 4: fronts along z, sparse-table discs in the plane (tiles of 64 x 32 with partial tiles, buckets of 32
    columns) -- the painted r^2 is the brute force's, bit for bit"""
    vol = shapes.random_blobs(shape, 5)
    rg, ridge, _ = _ridge_of(oracle, vol, inverse)
    _, want = oracle.paint(rg)
    got, flags = emu.paint_disc(ridge)
    assert flags == 0 and np.array_equal(got, want)
    if shape[0] > 8:                                   # a plane range, in chunks
        got, flags = emu.paint_disc(ridge, z_lo=3, z_hi=shape[0] - 2, chunk_planes=32)
        assert flags == 0 and np.array_equal(got, want[3:shape[0] - 2])


@pytest.mark.parametrize("variant", [1, 2, 3])
def test_painter_kernels_wide_items_and_doubled_pools(oracle, guard_behind, variant):
    vol = shapes.random_blobs((12, 14, 18), 9)
    rg, ridge, _ = _ridge_of(oracle, vol, True)
    _, want = oracle.paint(rg)
    got, flags, _ = emu.paint(ridge, variant=variant)
    assert flags == 0 and np.array_equal(got, want)


def test_painter_kernels_column_blocks(oracle, guard_behind):
    """stages 2-3 per block of 32 columns: 32 + 32 + 6"""
    vol = shapes.random_blobs((6, 9, 70), 4)
    rg, ridge, _ = _ridge_of(oracle, vol, True)
    _, want = oracle.paint(rg)
    got, flags, _ = emu.paint(ridge, xslab_cols=32)
    assert flags == 0 and np.array_equal(got, want)


@pytest.mark.parametrize("caps", [(2, 1), (3, 2)])
def test_painter_kernels_redo_paths(oracle, guard_behind, caps):
    """capacities of the fast kernels capped: lines are redone by the kernels with global scratch, same map"""
    from bonej2_b200.phantom import phantom_shapes
    vol = oracle.phantom_raster(phantom_shapes(24, 24, 24, 1), 24, 24, 24)
    for inverse in (False, True):
        rg, ridge, _ = _ridge_of(oracle, vol, inverse)
        _, want = oracle.paint(rg)
        got, flags, redo = emu.paint(ridge, cap=caps[0], fmax=caps[1])
        assert flags == 0 and np.array_equal(got, want)
        if inverse:
            assert redo > 10


@pytest.mark.parametrize("cuts,xcols", [((0, 9, 17), 0), ((0, 4, 8, 12, 17), 0), ((0, 6, 17), 32), ((0, 1, 2, 17), 0)])
@pytest.mark.parametrize("inverse", [False, True])
def test_painter_kernels_on_z_slabs(oracle, guard_behind, cuts, xcols, inverse):
    """the stepwise painter (x / y stages per slab, boundary staircases exported and imported, z sweeps) on slabs of
    one volume -- also slabs thinner than the largest ball and column blocks -- equals the whole-volume painting"""
    import math
    vol = shapes.random_blobs((17, 14, 40), 3)
    rg, ridge, _ = _ridge_of(oracle, vol, inverse)
    _, want = oracle.paint(rg)
    got, rc = emu.paint_slabs(ridge, cuts, math.isqrt(int(ridge.max())), xslab_cols=xcols)
    assert rc == 0 and np.array_equal(got, want)


def test_whole_chain_on_a_phantom(oracle, guard_behind):
    """EDT -> ridge -> paint, each kernel fed by the one before it"""
    from bonej2_b200.phantom import phantom_shapes
    vol = oracle.phantom_raster(phantom_shapes(32, 32, 32, 2), 32, 32, 32)
    for inverse in (False, True):
        o = oracle.process_image(vol, inverse=inverse, mask=False, taps=True)
        sq, mx = emu.edt_sq(vol, inverse=inverse)
        assert np.array_equal(sq, o["sq"])
        if mx == 0:
            continue
        rg, n = emu.ridge(sq)
        got, flags, _ = emu.paint(rg)
        _, want = oracle.paint(o["ridge"]) if "ridge" in o else (None, None)
        if want is not None:
            assert flags == 0 and np.array_equal(got, want)


@pytest.mark.parametrize("mask", [True, False])
@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("shape", [(9, 20, 32), (13, 11, 7), (5, 37, 33), (1, 9, 17), (2, 2, 2)])
def test_finish_kernel(oracle, guard, shape, inverse, mask):
    """2 sqrt, surface cleanup in the reference's visiting order, mask, NaN background, calibration and the statistics
    partials: the map within 1e-5 relative of the oracle's (float32, the north star's tolerance) with the same NaN
    pattern, the count exactly, sums to 1e-5"""
    vol = shapes.random_blobs(shape, 6)
    o = oracle.process_image(vol, inverse=inverse, mask=mask, pixel_width=0.5, taps=True)
    got, st = emu.finish(o["rsq"], vol if mask else None, inverse=inverse, pixel_width=0.5)
    want = o["map"]
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.allclose(got[ok], want[ok], rtol=1e-5, atol=0)
    assert st["count"] == o["stats"].count
    if st["count"]:
        assert st["sum"] == pytest.approx(o["stats"].sum, rel=1e-5) and st["sum2"] == pytest.approx(o["stats"].sum2, rel=1e-5)
        assert st["max"] == pytest.approx(o["stats"].max, rel=1e-5) and st["min"] == pytest.approx(o["stats"].min, rel=1e-5)


def test_finish_kernel_plane_range(oracle, guard):
    """output and statistics restricted to planes [z_lo, z_hi) of a slab with halo planes (bjt_dev_finish_range)"""
    vol = shapes.random_blobs((14, 12, 20), 8)
    o = oracle.process_image(vol, inverse=True, mask=True, taps=True)
    got, st = emu.finish(o["rsq"], vol, inverse=True, z_lo=3, z_hi=10)
    want = o["map"][3:10]
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.allclose(got[ok], want[ok], rtol=1e-5, atol=0)
    assert st["count"] == int(ok.sum())


@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("shape", [(9, 20, 32), (13, 11, 7), (1, 9, 17)])
def test_scatter_painter_kernels(oracle, guard_behind, shape, inverse):
    """the always-correct fallback (paint path 2): ridge compaction and the literal scatter of every ball"""
    vol = shapes.random_blobs(shape, 3)
    rg, ridge, _ = _ridge_of(oracle, vol, inverse)
    _, want = oracle.paint(rg)
    assert np.array_equal(emu.paint_scatter(ridge), want)


def test_phantom_kernel(oracle, guard):
    """the device rasteriser of the synthetic phantom equals the CPU one, whole and per z-slab"""
    from bonej2_b200.phantom import phantom_shapes
    sh = phantom_shapes(40, 36, 28, 5)
    want = oracle.phantom_raster(sh, 40, 36, 28)
    assert np.array_equal(emu.phantom(sh, 40, 36, 28), want)
    assert np.array_equal(emu.phantom(sh, 40, 36, 28, z0=9, dz=11), want[9:20])
