"""tests/golden/thickness_small.npz (made by tests/golden/make_golden.py): the oracle must still produce it (CPU), and the
CUDA path must reproduce it stage by stage through the C-ABI (GPU): squared EDT, ridge and painted r^2 bit for bit, maps
to 1e-5 relative with the same NaN pattern, statistics to 1e-5."""
import math
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "thickness_small.npz"))
NAMES = sorted({k.split("/")[0] for k in G.files})


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("inverse", [False, True])
def test_oracle_reproduces_fixture(oracle, name, inverse):
    tag = f"{name}/{'sp' if inverse else 'th'}"
    r = oracle.process_image(G[f"{name}/vol"], inverse=inverse, mask=True, pixel_width=0.5, taps=True)
    assert np.array_equal(r["sq"], G[f"{tag}/sq"])
    assert np.array_equal(np.where(r["ridge"] > 0, r["sq"], 0), G[f"{tag}/ridge"])
    assert np.array_equal(r["rsq"], G[f"{tag}/rsq"])
    assert np.array_equal(r["map"], G[f"{tag}/map"], equal_nan=True)
    st = r["stats"]
    assert np.array_equal(np.array([st.count, st.mean, st.sd, st.min, st.max]), G[f"{tag}/stats"], equal_nan=True)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("inverse", [False, True])
def test_cuda_reproduces_fixture(ctx, name, inverse):
    tag = f"{name}/{'sp' if inverse else 'th'}"
    vol = G[f"{name}/vol"]
    sq = ctx.edt_sq(vol, inverse=inverse)
    assert np.array_equal(sq, G[f"{tag}/sq"])
    ridge = ctx.ridge(sq)
    assert np.array_equal(ridge, G[f"{tag}/ridge"])
    for path in (0, 1, 2):
        assert np.array_equal(ctx.paint_rsq(ridge, path=path), G[f"{tag}/rsq"]), f"paint path {path}"
    m, st, _ = ctx.thickness(vol, inverse=inverse, mask=True, pixel_width=0.5)
    want = G[f"{tag}/map"]
    assert np.array_equal(np.isnan(m), np.isnan(want))
    ok = ~np.isnan(m)
    assert np.allclose(m[ok], want[ok], rtol=1e-5, atol=0)
    cnt, mean, sd, mn, mx = G[f"{tag}/stats"]
    assert st.count == int(cnt)
    if st.count:
        assert math.isclose(st.mean, mean, rel_tol=1e-5) and math.isclose(st.max, mx, rel_tol=1e-6)
        assert abs(st.sd - sd) <= 1e-5 * max(abs(sd), abs(mean) * 1e-3)
