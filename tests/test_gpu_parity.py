"""GPU parity: every stage of the CUDA path against the CPU oracle on the same inputs, through the
C-ABI (bonej2_b200._native binds include/bonej_thickness.h).  Integer stages are bit-exact; the
float map is compared at 1e-5 relative with an identical NaN pattern (BASELINE.json north_star)."""
import math

import numpy as np
import pytest

from tests import shapes

pytestmark = pytest.mark.gpu

CASES = shapes.edge_cases()
RTOL = 1e-5


def _phantom(oracle, n, seed=1):
    from bonej2_b200.phantom import phantom_shapes
    return oracle.phantom_raster(phantom_shapes(n, n, n, seed), n, n, n)


def _cmp_map(a, b):
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern differs"
    m = ~np.isnan(a)
    assert np.allclose(a[m], b[m], rtol=RTOL, atol=0)


def _cmp_stats(g, o):
    assert g.count == o.count
    if o.count == 0:
        assert math.isnan(g.mean)
        return
    for f in ("mean", "min", "max"):
        assert math.isclose(getattr(g, f), getattr(o, f), rel_tol=RTOL, abs_tol=1e-12), f
    # StackStatistics' one-pass sd = sqrt((n*sum2 - sum^2)/n/(n-1)) cancels catastrophically on a
    # (near-)constant map: its own rounding noise is ~ sqrt(n * 2^-53) * mean, so below that
    # level the reference value depends on its summation order.  Compare relative to the larger
    # of sd and that noise floor.
    noise = math.sqrt(max(o.count, 1) * 2.0 ** -52) * abs(o.mean)
    assert abs(g.sd - o.sd) <= RTOL * abs(o.sd) + 4 * noise, "sd"


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("inverse", [False, True])
def test_edt_bit_exact(ctx, oracle, name, inverse):
    vol = CASES[name]
    _, sq = oracle.edt(vol, inverse=inverse)
    assert np.array_equal(ctx.edt_sq(vol, inverse=inverse), sq)


def test_edt_threshold_semantics(ctx, oracle):
    """EDT_S1D thresholds raw bytes at 128 (the wrapper only admits 0/255, the stage itself does not care)"""
    rng = np.random.default_rng(0)
    vol = rng.integers(0, 256, (9, 33, 64), dtype=np.uint8)
    for inv in (False, True):
        assert np.array_equal(ctx.edt_sq(vol, inverse=inv), oracle.edt(vol, inverse=inv)[1])


def test_ridge_template(ctx, oracle):
    vals = np.array(sorted({a * a + b * b + c * c for a in range(12) for b in range(12) for c in range(12)} - {0}), np.int32)
    for g, o in zip(ctx.ridge_template(vals), oracle.ridge_template(vals)):
        assert np.array_equal(g, o)


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("inverse", [False, True])
def test_stages(ctx, oracle, name, inverse):
    vol = CASES[name]
    dist, sq = oracle.edt(vol, inverse=inverse)
    if sq.max() == 0:
        pytest.skip("no foreground")
    rg_o = oracle.ridge(dist)
    rg_g = ctx.ridge(sq)
    assert np.array_equal(rg_g > 0, rg_o > 0), "ridge set differs"
    assert np.array_equal(rg_g[rg_g > 0], sq[rg_g > 0])
    lt_o, rsq_o = oracle.paint(rg_o)
    for path in (2, 0, 1, 4, -1):
        assert np.array_equal(ctx.paint_rsq(rg_g, path=path), rsq_o), f"painted r^2 differs (path {path})"
    cl_o = oracle.cleanup(lt_o)
    cl_g, _ = ctx.cleanup(rsq_o, calibrate=False)
    assert np.array_equal(cl_g, cl_o), "cleanup differs"


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("mask", [False, True])
def test_pipeline(ctx, oracle, name, inverse, mask):
    vol = CASES[name]
    o = oracle.process_image(vol, inverse=inverse, mask=mask, pixel_width=0.37)
    m, st, tm = ctx.thickness(vol, inverse=inverse, mask=mask, pixel_width=0.37)
    _cmp_map(m, o["map"])
    _cmp_stats(st, o["stats"])
    assert tm.n_launches > 0


def test_wrapper_kat_2x2x2(ctx):
    """ThicknessWrapperTest.java:225-259"""
    v = np.zeros((2, 2, 2), np.uint8)
    (m_th, s_th, _), (m_sp, s_sp, _) = ctx.thickness_both(v, mask=False)
    assert s_th.count == 0 and np.isnan(m_th).all()
    assert s_sp.mean == 10.392304420471191 and s_sp.sd == 0.0 and s_sp.max == 10.392304420471191


def test_not_binary(ctx):
    from bonej2_b200._native import ThicknessError, BJT_E_NOT_BINARY
    v = np.zeros((4, 4, 32), np.uint8)
    v[1, 2, 3] = 7
    with pytest.raises(ThicknessError) as e:
        ctx.thickness(v)
    assert e.value.code == BJT_E_NOT_BINARY
    # every byte position of a volume whose size is not a multiple of the 16-byte vectors the check reads
    for pos in (0, 1, 15, 16, 17, 5 * 7 * 3 - 1, 5 * 7 * 3 - 2):
        v = np.full(5 * 7 * 3, 255, np.uint8)
        v[pos] = 254
        with pytest.raises(ThicknessError) as e:
            ctx.thickness(v.reshape(3, 7, 5))
        assert e.value.code == BJT_E_NOT_BINARY, pos
    ctx.thickness(np.full((3, 7, 5), 255, np.uint8))


def test_input_not_modified_and_slices(ctx, oracle):
    vol = CASES["blobs_40"]
    keep = vol.copy()
    outs, st, _ = ctx.thickness_slices([vol[z] for z in range(vol.shape[0])], inverse=False, mask=True)
    assert np.array_equal(vol, keep)
    o = oracle.process_image(vol, inverse=False, mask=True)
    _cmp_map(np.stack(outs), o["map"])
    _cmp_stats(st, o["stats"])


def test_phantom_bytes(ctx, oracle):
    from bonej2_b200.phantom import phantom_shapes
    s = phantom_shapes(96, 80, 64, seed=3)
    assert np.array_equal(ctx.phantom(s, 96, 80, 64), oracle.phantom_raster(s, 96, 80, 64))


@pytest.mark.parametrize("n,seed", [(64, 1), (128, 2)])
def test_phantom_pipeline(ctx, oracle, n, seed):
    vol = _phantom(oracle, n, seed)
    for inverse in (False, True):
        o = oracle.process_image(vol, inverse=inverse, mask=True, taps=True)
        assert np.array_equal(ctx.edt_sq(vol, inverse=inverse), o["sq"])
        m, st, tm = ctx.thickness(vol, inverse=inverse, mask=True)
        _cmp_map(m, o["map"])
        _cmp_stats(st, o["stats"])


def test_stats_kernel(ctx, oracle):
    rng = np.random.default_rng(1)
    m = rng.random((17, 33, 65)).astype(np.float32) * 40
    m[rng.random(m.shape) < 0.4] = np.nan
    _cmp_stats(ctx.stats(m), oracle.stats(m))


@pytest.mark.parametrize("caps", [(2, 1), (3, 2), (6, 3)])
def test_painter_redo_kernels(ctx, oracle, caps):
    """with the fast kernels' capacities capped, most lines go through the redo kernels (global scratch);
    the painted map must not change"""
    vol = _phantom(oracle, 64, 1)
    try:
        ctx.debug_set_paint_limits(*caps)
        for inverse in (False, True):
            dist, sq = oracle.edt(vol, inverse=inverse)
            rg = oracle.ridge(dist)
            _, rsq = oracle.paint(rg)
            ridge = np.where(rg > 0, sq, 0).astype(np.uint32)
            for path in (0, 1):
                assert np.array_equal(ctx.paint_rsq(ridge, path=path), rsq)
                assert ctx.debug_paint_redo_lines() > 50
    finally:
        ctx.debug_set_paint_limits(0, 0)


@pytest.mark.parametrize("shape", [(40, 36, 48), (33, 47, 64), (70, 35, 32), (128, 96, 64), (16, 200, 16)])
@pytest.mark.parametrize("inverse", [False, True])
def test_edt_odd_line_lengths(ctx, oracle, shape, inverse):
    """squared EDT, bit for bit, on line lengths that are not multiples of the kernels' tile sizes"""
    vol = shapes.random_blobs(shape, 11)
    _, want = oracle.edt(vol, inverse=inverse)
    assert np.array_equal(ctx.edt_sq(vol, inverse=inverse), want)


def test_disc_painter_integer_root(ctx):
    """the disc painter's correction-free integer square root is exact for every argument below 2^18"""
    assert ctx.debug_isqrt_check() == 0


@pytest.mark.parametrize("shape,inverse", [((96, 80, 72), True), ((80, 96, 40), False), ((130, 70, 33), True)])
def test_disc_painter_odd_shapes(ctx, oracle, shape, inverse):
    from bonej2_b200.phantom import phantom_shapes
    """paint path 4 (fronts along z + discs in the plane) on extents that are not multiples of its blocks
    (32 columns x 32 planes for the fronts, 64 x 32 tiles for the discs), bitwise against the oracle's scatter paint"""
    w, h, d = shape
    vol = oracle.phantom_raster(phantom_shapes(w, h, d, seed=3), w, h, d)
    dist, sq = oracle.edt(vol, inverse=inverse)
    rg = oracle.ridge(dist)
    _, rsq = oracle.paint(rg)
    ridge = np.where(rg > 0, sq, 0).astype(np.uint32)
    assert np.array_equal(ctx.paint_rsq(ridge, path=4), rsq)
