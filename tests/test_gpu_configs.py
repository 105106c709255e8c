"""BASELINE.json configurations.  configs[0] (256^3, Tb.Th + Tb.Sp, mask on) is compared with the CPU
oracle in full; at 512^3, where the brute-force oracle would take minutes, the CUDA path is checked
through size-independent properties: two independent painters agree bit for bit, two independent EDT
kernels agree, sqrt(EDT) is 1-Lipschitz, every foreground voxel of a run gets a value (count = number
of foreground voxels) and the two runs partition the volume."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _phantom(ctx, n, seed=1):
    from bonej2_b200.phantom import phantom_shapes
    return ctx.phantom(phantom_shapes(n, n, n, seed), n, n, n)


def test_config0_256_both_mask_on(ctx, oracle):
    vol = _phantom(ctx, 256)
    assert np.array_equal(vol, oracle.phantom_raster(__import__("bonej2_b200.phantom", fromlist=["x"]).phantom_shapes(256, 256, 256, 1), 256, 256, 256))
    (m_th, s_th, _), (m_sp, s_sp, _) = ctx.thickness_both(vol, mask=True)
    for inverse, m, st in ((False, m_th, s_th), (True, m_sp, s_sp)):
        o = oracle.process_image(vol, inverse=inverse, mask=True, taps=True)
        assert np.array_equal(ctx.edt_sq(vol, inverse=inverse), o["sq"]), "squared EDT not bit-exact"
        assert np.array_equal(np.isnan(m), np.isnan(o["map"]))
        ok = ~np.isnan(m)
        assert np.allclose(m[ok], o["map"][ok], rtol=1e-5, atol=0)
        assert st.count == o["stats"].count
        for f in ("mean", "sd", "max"):
            assert math.isclose(getattr(st, f), getattr(o["stats"], f), rel_tol=1e-5), f


@pytest.fixture(scope="module")
def vol512(ctx):
    return _phantom(ctx, 512)


@pytest.mark.parametrize("inverse", [False, True])
def test_512_properties(ctx, vol512, inverse):
    vol = vol512
    sq = ctx.edt_sq(vol, inverse=inverse)
    fg = (vol == 0) if inverse else (vol == 255)
    # background of the run is 0, foreground strictly positive
    assert np.array_equal(sq > 0, fg)
    # sqrt(EDT) is 1-Lipschitz along every axis: |sqrt(a) - sqrt(b)| <= 1  <=>  a <= b + 2 sqrt(b) + 1
    r = np.sqrt(sq.astype(np.float64))
    for ax in range(3):
        a = np.moveaxis(r, ax, 0)
        assert np.all(np.abs(a[1:] - a[:-1]) <= 1.0 + 1e-9)
    # an independent kernel (per-thread line scans, taken when the width is not a multiple of 4) agrees on a crop
    crop = np.ascontiguousarray(vol[:64, :70, :509])
    sq_crop_line = ctx.edt_sq(crop, inverse=inverse)
    crop4 = np.ascontiguousarray(vol[:64, :70, :512])
    # same planes, wider by 3 columns: the narrower crop can only have distances >= (fewer background voxels)
    assert np.all(sq_crop_line >= ctx.edt_sq(crop4, inverse=inverse)[:, :, :509])
    # two independent painters agree bit for bit
    ridge = ctx.ridge(sq)
    assert np.all(ridge[ridge > 0] == sq[ridge > 0]) and not np.any(ridge[~fg])
    sep = ctx.paint_rsq(ridge, path=4)
    assert np.array_equal(sep, ctx.paint_rsq(ridge, path=2)), "disc and scatter painters differ"
    assert np.array_equal(sep, ctx.paint_rsq(ridge, path=0)), "disc and sweep painters differ"
    # every foreground voxel is painted with at least its own squared distance
    assert np.all(sep[fg] >= sq[fg])
    del sep, ridge, r
    # full pipeline: count = number of foreground voxels of the run (mask on), max is the largest painted ball
    m, st, tm = ctx.thickness(vol, inverse=inverse, mask=True)
    assert st.count == int(fg.sum())
    assert np.array_equal(~np.isnan(m), fg)
    assert math.isclose(st.max, float(np.nanmax(m)), rel_tol=1e-6) and st.min >= 2.0
    assert tm.paint_path == 4


def test_512_line_kernel_equals_tile_kernel(ctx, vol512):
    """same volume, two shapes that take the two EDT code paths: pad the width to 516 (tile) vs 514 (line)"""
    base = vol512[:48, :48, :]
    a = np.zeros((48, 48, 516), np.uint8); a[:, :, :512] = base
    b = np.zeros((48, 48, 514), np.uint8); b[:, :, :512] = base
    for inverse in (False, True):
        sa, sb = ctx.edt_sq(a, inverse=inverse), ctx.edt_sq(b, inverse=inverse)
        if not inverse:
            assert np.array_equal(sa[:, :, :512], sb[:, :, :512])        # extra background columns beyond 512
        else:
            # inverted: the padding is foreground; compare where the padding cannot matter (distance to the
            # padding is at least 2, so any value <= 4 is decided by voxels both shapes share)
            sel = sa[:, :, :500] <= 4
            assert np.array_equal(sa[:, :, :500][sel], sb[:, :, :500][sel])


def test_config2_1024_device_resident_properties(ctx):
    """BASELINE.json configs[2] at one GPU (the bench's workload), device-resident like the bench: properties that
    need no oracle -- every voxel of the run's foreground gets a value (mask on), the two runs partition the
    volume, the largest thickness is the diameter of the largest inscribed ball (2 sqrt(max r^2), kept by the
    cleanup because that voxel is interior), every value is at least 2, and the run is repeatable bit for bit."""
    import torch
    from bonej2_b200.phantom import phantom_shapes
    n = 1024
    free, _ = torch.cuda.mem_get_info()
    if free < 40 * (1 << 30):
        pytest.skip("needs about 40 GB of free device memory")
    d_in = torch.empty(n ** 3, dtype=torch.uint8, device="cuda")
    ctx.dev_phantom(phantom_shapes(n, n, n, 1), n, n, n, 0, n, d_in.data_ptr())
    n_fg = int((d_in == 255).sum().item())
    assert int((d_in == 0).sum().item()) + n_fg == n ** 3
    d_out = torch.empty(n ** 3, dtype=torch.float32, device="cuda")
    counts = []
    for inverse in (False, True):
        st, tm = ctx.dev_thickness(d_in.data_ptr(), n, n, n, inverse=inverse, mask=True, d_out=d_out.data_ptr())
        counts.append(st.count)
        assert tm.paint_path == 4
        assert st.max == pytest.approx(2.0 * math.sqrt(tm.rsq_max), rel=1e-6)
        assert st.min >= 2.0
        assert int(torch.isnan(d_out).sum().item()) == n ** 3 - st.count
        first = d_out[::4099].clone()
        st2, _ = ctx.dev_thickness(d_in.data_ptr(), n, n, n, inverse=inverse, mask=True, d_out=d_out.data_ptr())
        assert (st2.count, st2.mean, st2.sd, st2.max) == (st.count, st.mean, st.sd, st.max)
        assert torch.equal(torch.nan_to_num(first, nan=-1.0), torch.nan_to_num(d_out[::4099], nan=-1.0))
    assert counts == [n_fg, n ** 3 - n_fg]


@pytest.mark.parametrize("n", [512, 1024])
@pytest.mark.parametrize("inverse", [False, True])
def test_squared_edt_bit_exact_at_bench_sizes(ctx, oracle, n, inverse):
    """the sizes the benchmark runs (BASELINE.json configs[1], configs[2]): the CUDA squared EDT against an
    independent exact algorithm (oracle/edt_exact.c, pinned to the restatement in tests/test_oracle_golden.py), every
    voxel, both runs.  Lines of 1024 with 16-line tiles and the long outward walks of the background run are the
    paths small cases hardly reach."""
    import torch
    if n == 1024 and torch.cuda.mem_get_info()[0] < 24 * (1 << 30):
        pytest.skip("needs about 24 GB of free device memory")
    vol = _phantom(ctx, n)
    got = ctx.edt_sq(vol, inverse=inverse)
    want = oracle.edt_exact_sq(vol, inverse=inverse)
    assert got.max() == want.max()
    for z0 in range(0, n, 64):                                     # bounded temporaries
        assert np.array_equal(got[z0:z0 + 64], want[z0:z0 + 64]), (n, inverse, z0)
    if n == 1024:
        ctx.release_workspace()


@pytest.mark.parametrize("shape", [(2048, 32, 64), (32, 2048, 64), (300, 700, 96), (64, 64, 1344)])
def test_squared_edt_pair_kernel_long_lines_and_redo(ctx, oracle, shape):
    """edt_tile16_kernel off the beaten track: lines of 2048 (16-column tiles, the shape of the 2048^3 slabs), widths
    that are multiples of 16 or 32 only, and -- on a volume with a handful of background voxels -- walks that leave the
    pad rows everywhere and squared distances far above the 16-bit clamp (the in-kernel 32-bit redo), next to a dense
    volume on which nothing saturates"""
    rng = np.random.default_rng(5)
    dense = (rng.random(shape) < 0.7).astype(np.uint8) * 255
    sparse = np.full(shape, 255, np.uint8)
    idx = tuple(rng.integers(0, s, 3) for s in shape)
    sparse[idx] = 0
    half = sparse.copy()                                              # one half dense, the other nearly empty: both paths in one launch
    half[:, :, : shape[2] // 2] = dense[:, :, : shape[2] // 2]
    for vol in (dense, sparse, half):
        for inverse in (False, True):
            got = ctx.edt_sq(vol, inverse=inverse)
            want = oracle.edt_exact_sq(vol, inverse=inverse)
            assert np.array_equal(got, want), (shape, inverse)
    assert oracle.edt_exact_sq(sparse).max() > 55000


def test_1024_disc_painter_equals_sweep_painter_on_the_device(ctx):
    """BASELINE.json configs[2], background run (the large balls): the default painter (path 4, fronts along z + discs
    in the plane) against the separable sweeps (path 0, a different algorithm, oracle-checked on small volumes), bit
    for bit, device-resident"""
    import torch
    from bonej2_b200.phantom import phantom_shapes
    n = 1024
    if torch.cuda.mem_get_info()[0] < 100 * (1 << 30):
        pytest.skip("needs about 100 GB of free device memory (the sweep painter's lists)")
    N = n ** 3
    d_in = torch.empty(N, dtype=torch.uint8, device="cuda")
    ctx.dev_phantom(phantom_shapes(n, n, n, 1), n, n, n, 0, n, d_in.data_ptr())
    gx = torch.empty(N, dtype=torch.int16, device="cuda")
    a = torch.empty(N, dtype=torch.int32, device="cuda")
    sq = torch.empty(N, dtype=torch.int32, device="cuda")
    ctx.dev_edt_x(d_in.data_ptr(), n, n, n, True, 128, gx.data_ptr())
    ctx.dev_edt_y(gx.data_ptr(), n, n, n, n, a.data_ptr())
    ctx.dev_edt_z(a.data_ptr(), n, n, n, n, sq.data_ptr())
    del gx
    ctx.dev_ridge(sq.data_ptr(), n, n, n, a.data_ptr())            # a = ridge
    ctx.dev_paint(a.data_ptr(), n, n, n, sq.data_ptr(), path=4)    # sq = painted r^2, path 4
    b = torch.empty(N, dtype=torch.int32, device="cuda")
    ctx.dev_paint(a.data_ptr(), n, n, n, b.data_ptr(), path=0)
    torch.cuda.synchronize()
    same = torch.equal(sq, b)
    del a, b, sq, d_in
    ctx.release_workspace()
    assert same
