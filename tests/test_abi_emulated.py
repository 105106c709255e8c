"""The `-m gpu` parity tests of the C-ABI, run where there is no GPU: bonej2_b200._native is bound to the library's
own sources compiled against the CPU emulation of CUDA (tests/emu_native.py), and the test functions of
tests/test_gpu_parity.py / test_golden_fixtures.py are called as they are.  This exercises api.cu's orchestration
(workspace, readbacks, fast paths for empty / full volumes, paint path selection, slices entry point, "Both") together
with every kernel of the Thickness pipeline.  It is a logic check; the GPU run remains the parity gate."""
import numpy as np
import pytest

from tests import shapes
from tests import test_gpu_parity as G
from tests import test_golden_fixtures as F

SMALL = [n for n in sorted(G.CASES) if G.CASES[n].size <= 70000]


@pytest.fixture(scope="module")
def emu_ctx():
    from bonej2_b200 import _native
    from tests.emu_native import emulated_library
    with emulated_library():
        c = _native.Context(0)
        assert "emulation" in c.device_info()["name"]
        yield c
        c.close()


# both runs for the small cases, the background run (larger balls, longer lists) alone for the bigger ones
STAGE_CASES = [(n, inv) for n in SMALL for inv in (False, True) if inv or G.CASES[n].size <= 20000]


@pytest.mark.parametrize("name,inverse", STAGE_CASES)
def test_stages(emu_ctx, oracle, name, inverse):
    G.test_stages(emu_ctx, oracle, name, inverse)


@pytest.mark.parametrize("name", ["blobs_40", "noncubic_37x101x13", "single_bg_voxel", "sphere_r6", "2x2x2_full", "row_1d"])
@pytest.mark.parametrize("inverse", [False, True])
def test_pipeline_mask_on(emu_ctx, oracle, name, inverse):
    G.test_pipeline(emu_ctx, oracle, name, inverse, True)


@pytest.mark.parametrize("name", ["blobs_40", "2x2x2_zero", "single_bg_voxel"])
def test_pipeline_mask_off(emu_ctx, oracle, name):
    G.test_pipeline(emu_ctx, oracle, name, True, False)


def test_small_entry_points(emu_ctx, oracle):
    G.test_edt_threshold_semantics(emu_ctx, oracle)
    G.test_ridge_template(emu_ctx, oracle)
    G.test_wrapper_kat_2x2x2(emu_ctx)
    G.test_not_binary(emu_ctx)
    G.test_input_not_modified_and_slices(emu_ctx, oracle)
    G.test_phantom_bytes(emu_ctx, oracle)
    G.test_stats_kernel(emu_ctx, oracle)


@pytest.mark.parametrize("name", F.NAMES)
@pytest.mark.parametrize("inverse", [False, True])
def test_golden_fixtures(emu_ctx, name, inverse):
    F.test_cuda_reproduces_fixture(emu_ctx, name, inverse)


def test_both_maps_in_one_call(emu_ctx, oracle):
    """bjt_thickness_both_u8: one upload, Tb.Th then Tb.Sp"""
    vol = shapes.random_blobs((14, 12, 32), 6)
    (m_th, s_th, _), (m_sp, s_sp, _) = emu_ctx.thickness_both(vol, mask=True)
    for inverse, m, st in ((False, m_th, s_th), (True, m_sp, s_sp)):
        o = oracle.process_image(vol, inverse=inverse, mask=True)
        G._cmp_map(m, o["map"])
        G._cmp_stats(st, o["stats"])


@pytest.mark.parametrize("depth", [75, 33])
def test_both_maps_streamed_in_chunks(emu_ctx, oracle, depth):
    """host-buffer calls paint, finish and send the map home in chunks of planes that halve towards the end of a run
    (75 planes: 64 + 11; 33: 32 + 1 -- a one-plane last chunk, whose cleanup reads planes of the chunk before)"""
    vol = shapes.random_blobs((depth, 9, 32), 4)
    (m_th, s_th, _), (m_sp, s_sp, _) = emu_ctx.thickness_both(vol, mask=True)
    for inverse, m, st in ((False, m_th, s_th), (True, m_sp, s_sp)):
        o = oracle.process_image(vol, inverse=inverse, mask=True)
        G._cmp_map(m, o["map"])
        G._cmp_stats(st, o["stats"])


def test_thickness_wrapper_command(emu_ctx):
    """the C++ mirror of the ThicknessWrapper command (tests/test_thickness_wrapper.py, the tests that produce maps)"""
    from bonej2_b200 import wrapper
    from tests import test_thickness_wrapper as W
    calls = [W.test_anisotropy_headless_warns_and_continues, W.test_bad_map_choice_throws, W.test_results_kat,
             W.test_default_unit_and_rows_accumulate]
    calls += [lambda c=c, th=th, sp=sp: W.test_map_outputs(c, th, sp)
              for c, th, sp in (("Trabecular thickness", True, False), ("Trabecular separation", False, True), ("Both", True, True))]
    for call in calls:
        wrapper.shared_table_reset()         # the results table is process-wide, like the reference's SharedTable
        call()


@pytest.mark.parametrize("shape", [(40, 36, 48), (33, 47, 64), (16, 200, 16)])
def test_edt_odd_line_lengths(emu_ctx, oracle, shape):
    G.test_edt_odd_line_lengths(emu_ctx, oracle, shape, True)


def test_painter_redo_kernels_through_the_abi(emu_ctx, oracle):
    """bjt_debug_set_paint_limits + bjt_paint_rsq on both item widths, redo-line counter read back"""
    from bonej2_b200.phantom import phantom_shapes
    vol = oracle.phantom_raster(phantom_shapes(24, 24, 24, 1), 24, 24, 24)
    try:
        emu_ctx.debug_set_paint_limits(2, 1)
        dist, sq = oracle.edt(vol, inverse=True)
        rg = oracle.ridge(dist)
        _, rsq = oracle.paint(rg)
        ridge = np.where(rg > 0, sq, 0).astype(np.uint32)
        for path in (0, 1, 2):
            assert np.array_equal(emu_ctx.paint_rsq(ridge, path=path), rsq)
            if path < 2:
                assert emu_ctx.debug_paint_redo_lines() > 10
    finally:
        emu_ctx.debug_set_paint_limits(0, 0)


def test_graft_entry_smoke(emu_ctx):
    """__graft_entry__.smoke() -- what the driver runs on the GPU box before the bench -- on the emulated library"""
    import __graft_entry__ as g
    g.smoke()


def test_hyperstack_and_raw_file_front_ends(emu_ctx, oracle, tmp_path):
    """the callers either side of the path: per-subspace batch over a hyperstack, raw file to raw file
    through page-locked buffers"""
    from tests import test_hyperstack as H
    from tests import test_rawio as R
    H.test_batched_thickness_matches_oracle(oracle)
    R.test_thickness_raw_file_to_file(tmp_path, oracle, "Both")
    R.test_thickness_raw_refuses_non_binary(tmp_path)


def test_consumers_of_the_map(emu_ctx):
    """segmented statistics (segstats.cu: __match_any_sync peer groups, two passes) through the C-ABI"""
    from tests import test_gpu_consumers as K
    K.test_label_stats_match_get_mean_std_dev(emu_ctx, 1)
    K.test_label_stats_rejects_unknown_label(emu_ctx)
    K.test_slice_stats_match_slice_geometry(emu_ctx, (5, 3, 20, 17))
    K.test_slice_stats_match_slice_geometry(emu_ctx, None)
    K.test_slice_stats_bad_roi(emu_ctx)


def test_find_ridge_points(emu_ctx):
    """FindRidgePoints (ridgepoints.cu on the CUDA EDT) through the C-ABI"""
    from tests import test_ridge_points as P
    P.test_cuda_matches_restatement(emu_ctx, 1, 0.6)
    P.test_cuda_matches_restatement(emu_ctx, 2, 0.2)
    P.test_cuda_degenerate(emu_ctx)
