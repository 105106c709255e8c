"""N > 1 path on CPU: two processes over gloo run bonej2_b200.sharded.thickness_slab with the
oracle-backed stage backend; the assembled result must equal the oracle on the whole volume (maps
bitwise, statistics to 1e-12).  Exercises the z<->y transposes, the r_max halo gather with clipping,
the no-background / no-foreground shortcuts and the statistics combination."""
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


def _thicker_slab_narrower_blocks(dz, h, w):
    """stand-in for the library's rule (column blocks sized by the slab's volume) at test scale: ranks whose slabs
    differ in thickness prefer different widths, as they do at 2048^3 on 8 GPUs, or at 1024^3 on 2 once the painter's
    split is cost-balanced"""
    return max(3, 40 // max(dz, 1))


def _worker(rank, world, port, vol_path, out_dir, inverse, mask, paint_mode, kcap=None, blocks=False):
    sys.path.insert(0, ROOT)
    from bonej2_b200 import sharded
    from tests.sharded_cpu_backend import OracleStages
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    vol = np.load(vol_path)
    d, h, w = vol.shape
    zs = sharded.split_ranges(d, world)
    slab = torch.from_numpy(vol[zs[rank][0]:zs[rank][1]].copy())
    comm = sharded.Comm()
    kw = {} if kcap is None else {"boundary_kcap": kcap}
    stages = OracleStages()
    if blocks:
        stages.block_rule = _thicker_slab_narrower_blocks
    out, stats = sharded.thickness_slab(stages, comm, slab, (w, h, d), inverse, mask, 0.5, paint_mode=paint_mode, **kw)
    np.save(os.path.join(out_dir, f"map_{rank}.npy"), out.numpy())
    if rank == 0:
        np.save(os.path.join(out_dir, "stats.npy"), np.array([stats["count"], stats["mean"], stats["sd"], stats["min"], stats["max"]]))
    dist.destroy_process_group()


def _run(vol, inverse, mask, tmp_path, world=2, paint_mode="exchange", kcap=None, blocks=False):
    vol_path = os.path.join(tmp_path, "vol.npy")
    np.save(vol_path, vol)
    port = _free_port()
    mp.spawn(_worker, args=(world, port, vol_path, str(tmp_path), inverse, mask, paint_mode, kcap, blocks), nprocs=world, join=True)
    maps = [np.load(os.path.join(tmp_path, f"map_{r}.npy")) for r in range(world)]
    return np.concatenate(maps, axis=0), np.load(os.path.join(tmp_path, "stats.npy"))


def _check(vol, inverse, mask, tmp_path, world=2, paint_mode="exchange", kcap=None, blocks=False):
    from oracle import lt_oracle
    got, st = _run(vol, inverse, mask, tmp_path, world, paint_mode, kcap, blocks)
    ref = lt_oracle.process_image(vol, inverse=inverse, mask=mask, pixel_width=0.5)
    assert np.array_equal(np.isnan(got), np.isnan(ref["map"]))
    ok = ~np.isnan(got)
    assert np.array_equal(got[ok], ref["map"][ok])
    assert int(st[0]) == ref["stats"].count
    if ref["stats"].count:
        assert math.isclose(st[1], ref["stats"].mean, rel_tol=1e-12)
        assert math.isclose(st[2], ref["stats"].sd, rel_tol=1e-9, abs_tol=1e-9)
        assert st[4] == ref["stats"].max


@pytest.mark.parametrize("paint_mode", ["exchange", "halo"])
@pytest.mark.parametrize("inverse", [False, True])
def test_two_ranks_blobs(tmp_path, inverse, paint_mode):
    from tests import shapes
    _check(shapes.random_blobs((17, 14, 19), 3), inverse, True, str(tmp_path), paint_mode=paint_mode)


def test_four_ranks_thin_slabs_exchange(tmp_path):
    """slabs thinner than the largest ball: staircases cross more than one slab face (several sources per side)"""
    vol = np.zeros((16, 14, 14), np.uint8)
    vol[1:15, 1:13, 1:13] = 255                                         # r_max = 6 > slab thickness 4
    vol[7, 6, 6] = 0
    _check(vol, False, True, str(tmp_path), world=4, paint_mode="exchange")


def test_boundary_overflow_falls_back_to_halo_repaint(tmp_path):
    """one staircase step per column is not enough here: every rank must notice and repaint slab + halo"""
    from tests import shapes
    vol = shapes.random_blobs((17, 14, 19), 3)
    _check(vol, True, True, str(tmp_path), paint_mode="auto", kcap=1)
    with pytest.raises(Exception):
        _check(vol, True, True, str(tmp_path), paint_mode="exchange", kcap=1)       # no fallback allowed: an error, not a wrong map


def test_empty_slabs_exchange(tmp_path):
    """more ranks than planes: ranks with no plane take part in the collectives only"""
    from tests import shapes
    _check(shapes.random_blobs((3, 9, 8), 2), True, False, str(tmp_path), world=4, paint_mode="exchange")


def test_two_ranks_big_radius_halo_clipped(tmp_path):
    """r_max larger than a slab: the halo gather reaches across the whole volume"""
    vol = np.zeros((12, 10, 10), np.uint8)
    vol[:, 1:9, 1:9] = 255
    vol[0] = 0
    _check(vol, False, False, str(tmp_path), paint_mode="halo")
    _check(vol, False, False, str(tmp_path), paint_mode="exchange")


def test_two_ranks_no_background(tmp_path):
    _check(np.full((4, 3, 5), 255, np.uint8), False, False, str(tmp_path))


def test_two_ranks_no_foreground(tmp_path):
    _check(np.zeros((4, 3, 5), np.uint8), False, True, str(tmp_path))


def test_three_ranks_uneven(tmp_path):
    from tests import shapes
    _check(shapes.random_blobs((11, 7, 9), 5), True, True, str(tmp_path), world=3)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("inverse", [False, True])
def test_two_ranks_column_blocks_agreed(tmp_path, inverse):
    """the ranks' slabs differ in thickness (9 / 8 planes, and the cost-balanced split of the painter moves more), so
    each would cut the width into different column blocks; staircases are exchanged per block, so they have to
    agree first (sharded.paint_exchange) -- without that the exchange does not even terminate"""
    from tests import shapes
    _check(shapes.random_blobs((17, 14, 19), 3), inverse, True, str(tmp_path), paint_mode="exchange", blocks=True)


@pytest.mark.timeout(300)
def test_three_ranks_column_blocks_and_an_empty_slab(tmp_path):
    from tests import shapes
    _check(shapes.random_blobs((11, 7, 9), 5), True, True, str(tmp_path), world=3, blocks=True)
    _check(shapes.random_blobs((2, 9, 8), 2), True, False, str(tmp_path), world=3, paint_mode="exchange", blocks=True)


def test_even_block_cols():
    from bonej2_b200.sharded import even_block_cols
    assert even_block_cols(1024, 896) == 512
    assert even_block_cols(1024, 1024) == 1024 and even_block_cols(1024, 1 << 20) == 1 << 20
    assert even_block_cols(2048, 768) == 768           # 3 blocks: 683 -> 768 (the quantum), not above the limit
    assert even_block_cols(2048, 896) == 768
    assert even_block_cols(19, 5) == 5                  # test-scale widths: the limit itself
    for w in (1000, 1024, 2048, 4096, 8192):
        for limit in range(512, 4097, 128):
            c = even_block_cols(w, limit)
            assert c <= limit and c % 32 == 0 and (c >= w or -(-w // c) == -(-w // limit))


def test_split_ranges():
    from bonej2_b200.sharded import split_ranges
    assert split_ranges(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert split_ranges(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]


def test_balanced_ranges():
    from bonej2_b200.sharded import balanced_ranges
    assert balanced_ranges([1] * 10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert balanced_ranges([10, 1, 1, 1, 1, 1, 1, 1, 1, 10], 2) == [(0, 5), (5, 10)]
    assert balanced_ranges([0, 0, 5, 0, 0], 4) == [(0, 3), (3, 3), (3, 3), (3, 5)]          # empty ranges are legal
    assert balanced_ranges([], 3) == [(0, 0)] * 3
    for w, g in (([3, 1, 4, 1, 5, 9, 2, 6], 3), ([1, 1], 4), ([7], 2)):
        r = balanced_ranges(w, g)
        assert len(r) == g and r[0][0] == 0 and r[-1][1] == len(w)
        assert all(a <= b for a, b in r) and all(r[i][1] == r[i + 1][0] for i in range(g - 1))
