"""The product's multi-GPU stage code on the CPU: bonej2_b200.sharded.CudaStages -- the class that drives the CUDA
kernels through the C-ABI on a GPU box -- runs here unchanged on CPU tensors, with the kernels served by their
CPU emulation (tests/kernels_emu.EmuContext), two to four ranks over gloo.  Together with the oracle-backed backend
of tests/test_sharded_cpu.py this covers what a multi-GPU run does except NCCL and the hardware: slab-local EDT
passes on u16 / u32 buffers, transposes, the halo gather, the ridge on fetched planes, the disc painter on a plane
range of slab + halo, the finish kernel on a plane range with a shifted original pointer, statistics combination."""
import math
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, vol_path, out_dir, inverse, mask):
    sys.path.insert(0, ROOT)
    from bonej2_b200 import sharded
    from tests.kernels_emu import EmuContext
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    vol = np.load(vol_path)
    d, h, w = vol.shape
    zs = sharded.split_ranges(d, world)
    slab = torch.from_numpy(vol[zs[rank][0]:zs[rank][1]].copy())
    stages = sharded.CudaStages(EmuContext(), torch.device("cpu"))
    out, stats = sharded.thickness_slab(stages, sharded.Comm(), slab, (w, h, d), inverse, mask, 0.5)
    np.save(os.path.join(out_dir, f"map_{rank}.npy"), out.numpy())
    if rank == 0:
        np.save(os.path.join(out_dir, "stats.npy"), np.array([stats["count"], stats["mean"], stats["sd"], stats["min"], stats["max"]]))
    dist.destroy_process_group()


def _check(vol, inverse, mask, tmp_path, world=2):
    from oracle import lt_oracle
    vol_path = os.path.join(tmp_path, "vol.npy")
    np.save(vol_path, vol)
    mp.spawn(_worker, args=(world, _free_port(), vol_path, str(tmp_path), inverse, mask), nprocs=world, join=True)
    got = np.concatenate([np.load(os.path.join(tmp_path, f"map_{r}.npy")) for r in range(world)], axis=0)
    st = np.load(os.path.join(tmp_path, "stats.npy"))
    ref = lt_oracle.process_image(vol, inverse=inverse, mask=mask, pixel_width=0.5)
    assert np.array_equal(np.isnan(got), np.isnan(ref["map"]))
    ok = ~np.isnan(got)
    assert np.allclose(got[ok], ref["map"][ok], rtol=1e-5, atol=0)
    assert int(st[0]) == ref["stats"].count
    if ref["stats"].count:
        assert math.isclose(st[1], ref["stats"].mean, rel_tol=1e-5) and math.isclose(st[4], ref["stats"].max, rel_tol=1e-5)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("inverse", [False, True])
def test_two_ranks_product_stages(tmp_path, inverse):
    from tests import shapes
    _check(shapes.random_blobs((18, 14, 40), 3), inverse, True, str(tmp_path))


@pytest.mark.timeout(900)
def test_three_ranks_product_stages_mask_off(tmp_path):
    from tests import shapes
    _check(shapes.random_blobs((13, 10, 32), 7), True, False, str(tmp_path), world=3)


@pytest.mark.parametrize("inverse", [False, True])
def test_single_rank_product_stages(oracle, inverse):
    """world size 1, in process (what tests/test_gpu_sharded.py::test_single_rank_matches_oracle does on a GPU)"""
    from bonej2_b200 import sharded
    from tests import shapes
    from tests.kernels_emu import EmuContext
    vol = shapes.random_blobs((11, 13, 24), 2)
    d, h, w = vol.shape
    stages = sharded.CudaStages(EmuContext(), torch.device("cpu"))
    out, st = sharded.thickness_slab(stages, sharded.Comm(), torch.from_numpy(vol.copy()), (w, h, d), inverse, True, 0.7)
    ref = oracle.process_image(vol, inverse=inverse, mask=True, pixel_width=0.7)
    got = out.numpy()
    assert np.array_equal(np.isnan(got), np.isnan(ref["map"]))
    ok = ~np.isnan(got)
    assert np.allclose(got[ok], ref["map"][ok], rtol=1e-5, atol=0)
    assert st["count"] == ref["stats"].count


@pytest.mark.timeout(900)
def test_four_ranks_thin_slabs_product_stages(tmp_path):
    """slabs thinner than the largest ball: halos reach across several slabs, the painter's plane range sits in the
    middle of a much thicker ridge volume"""
    vol = np.zeros((16, 14, 32), np.uint8)
    vol[1:15, 1:13, 1:31] = 255                                         # r_max = 6 > slab thickness 4
    vol[7, 6, 6] = 0
    _check(vol, False, True, str(tmp_path), world=4)


@pytest.mark.parametrize("inverse", [False, True])
def test_chunked_run_with_a_sink_equals_the_plain_run(inverse):
    """thickness_slab(..., sink=...) paints and finishes the slab in chunks of planes and hands every finished chunk to
    the sink in order (what the bench's end-to-end loop uses to send planes home behind the painter): the map, the
    statistics and the pieces the sink saw are those of the plain run"""
    from bonej2_b200 import sharded
    from tests import shapes
    from tests.kernels_emu import EmuContext
    vol = shapes.random_blobs((75, 9, 32), 4)                              # 75 planes: chunks of 32, 32, 11
    d, h, w = vol.shape
    stages = sharded.CudaStages(EmuContext(), torch.device("cpu"))
    slab = torch.from_numpy(vol.copy())
    plain, st0 = sharded.thickness_slab(stages, sharded.Comm(), slab, (w, h, d), inverse, True, 0.5)
    seen, pieces = [], torch.full((d, h, w), -7.0)
    def sink(a, b, view):
        seen.append((a, b))
        pieces[a:b] = view
    out, st1 = sharded.thickness_slab(stages, sharded.Comm(), slab, (w, h, d), inverse, True, 0.5, None, None, sink)
    assert len(seen) >= 2 and seen[0][0] == 0 and seen[-1][1] == d and all(x[1] == y[0] for x, y in zip(seen, seen[1:]))
    for t in (out, pieces):
        assert np.array_equal(np.nan_to_num(t.numpy(), nan=-1.0), np.nan_to_num(plain.numpy(), nan=-1.0))
    assert st1["count"] == st0["count"] and st1["max"] == st0["max"] and st1["min"] == st0["min"]
    assert st1["mean"] == pytest.approx(st0["mean"], rel=1e-12) and st1["sd"] == pytest.approx(st0["sd"], rel=1e-9)
