"""bjt_create_multi: the thickness calls of the C-ABI on a context over several devices (include/bonej_thickness.h).
The stack is cut into z-slabs inside the library, one host thread per device, lines and halos exchanged through peer
memory; the result must be the single-device map bit for bit (and therefore the oracle's within the float tolerance).
This is the path the one blocking call of the reference seam (ThicknessWrapper.java:127-134, 190-196) takes to reach
several GPUs.  Devices may be named more than once, so the tests run on one GPU (`-m gpu`) and, at small sizes, on
the CPU emulation of CUDA (tests/cuda_emu) where there is none; tests/test_gpu_sharded.py holds the runs on distinct GPUs."""
import numpy as np
import pytest

from tests import shapes
from tests import test_gpu_parity as G


@pytest.fixture(scope="module", params=[pytest.param("gpu", marks=pytest.mark.gpu), "emulated"])
def lib_kind(request):
    if request.param == "emulated":
        from tests.emu_native import emulated_library
        with emulated_library():
            yield "emulated"
    else:
        yield "gpu"


@pytest.fixture(scope="module")
def one(lib_kind):
    from bonej2_b200 import _native
    c = _native.Context(0)
    yield c
    c.close()


def _multi(nslabs, kind):
    """slabs dealt round-robin over the GPUs of the box (one GPU: they share it); the emulation has one device"""
    from bonej2_b200 import _native
    import torch
    ndev = torch.cuda.device_count() if (kind == "gpu" and torch.cuda.is_available()) else 1
    return _native.Context(devices=[i % max(ndev, 1) for i in range(nslabs)])


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("nslabs", [2, 3, 5])
@pytest.mark.parametrize("shape,seed", [((24, 20, 28), 3), ((9, 33, 17), 8)])
def test_both_equals_single_device(lib_kind, one, oracle, nslabs, shape, seed):
    vol = shapes.random_blobs(shape, seed)
    with _multi(nslabs, lib_kind) as m:
        assert m.device_count() == nslabs
        (a_th, sa_th, t_th), (a_sp, sa_sp, _) = m.thickness_both(vol, mask=True, pixel_width=0.5)
    (b_th, sb_th, _), (b_sp, sb_sp, _) = one.thickness_both(vol, mask=True, pixel_width=0.5)
    assert np.array_equal(_bits(a_th), _bits(b_th)) and np.array_equal(_bits(a_sp), _bits(b_sp))
    for sa, sb in ((sa_th, sb_th), (sa_sp, sb_sp)):
        assert sa.count == sb.count and sa.min == sb.min and sa.max == sb.max
        assert sa.mean == pytest.approx(sb.mean, rel=1e-12) and sa.sd == pytest.approx(sb.sd, rel=1e-9)
    assert t_th.n_launches > 0
    o = oracle.process_image(vol, inverse=True, mask=True, pixel_width=0.5)
    G._cmp_map(a_sp, o["map"])
    G._cmp_stats(sa_sp, o["stats"])


@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("mask", [False, True])
def test_single_run_and_slices(lib_kind, one, inverse, mask):
    vol = shapes.random_blobs((17, 22, 26), 11)
    with _multi(4, lib_kind) as m:
        a, sa, _ = m.thickness(vol, inverse=inverse, mask=mask)
        outs, ss, _ = m.thickness_slices([vol[z] for z in range(vol.shape[0])], inverse=inverse, mask=mask)
        _, s_only, _ = m.thickness(vol, inverse=inverse, mask=mask, want_map=False)
    b, sb, _ = one.thickness(vol, inverse=inverse, mask=mask)
    assert np.array_equal(_bits(a), _bits(b)) and np.array_equal(_bits(np.stack(outs)), _bits(b))
    assert sa.count == sb.count == ss.count == s_only.count and sa.max == sb.max
    assert s_only.mean == pytest.approx(sb.mean, rel=1e-12)


@pytest.mark.parametrize("name", ["2x2x2_zero", "2x2x2_full", "single_bg_voxel", "row_1d", "sphere_r6"])
def test_edge_volumes(lib_kind, one, oracle, name):
    """empty and full volumes (constant fast paths), fewer planes than devices, one plane"""
    vol = G.CASES[name]
    with _multi(3, lib_kind) as m:
        for inverse in (False, True):
            a, sa, _ = m.thickness(vol, inverse=inverse, mask=True)
            b, sb, _ = one.thickness(vol, inverse=inverse, mask=True)
            assert np.array_equal(_bits(a), _bits(b)), (name, inverse)
            assert sa.count == sb.count
            if sb.count:
                assert sa.max == sb.max and sa.mean == pytest.approx(sb.mean, rel=1e-12)
            else:
                assert np.isnan(sa.mean) and np.isnan(sb.mean)


def test_not_binary_and_bad_arguments(lib_kind):
    from bonej2_b200 import _native
    vol = np.zeros((6, 5, 4), np.uint8)
    vol[5, 4, 3] = 9                                    # in the last slab only
    with _multi(3, lib_kind) as m:
        with pytest.raises(_native.ThicknessError) as e:
            m.thickness(vol)
        assert e.value.code == _native.BJT_E_NOT_BINARY
        vol[5, 4, 3] = 255
        a, _, _ = m.thickness(vol, inverse=True)        # and the context is still usable
        assert np.isnan(a[5, 4, 3])
    with pytest.raises(_native.ThicknessError):
        _native.Context(devices=[])
    with pytest.raises(_native.ThicknessError):
        _native.Context(devices=[0, 4096])


def test_thin_slabs_far_reach(lib_kind, one):
    """balls much thicker than a slab: the halo of a slab spans several other slabs"""
    vol = np.full((30, 26, 26), 255, np.uint8)
    vol[0, 0, 0] = 0
    vol[29, 25, 25] = 0
    with _multi(6, lib_kind) as m:
        a, sa, tm = m.thickness(vol, inverse=False, mask=True)
    b, sb, _ = one.thickness(vol, inverse=False, mask=True)
    assert tm.rsq_max > 5 * 5 * 3
    assert np.array_equal(_bits(a), _bits(b)) and sa.count == sb.count and sa.max == sb.max
