"""N > 1 path on CPU: two processes over gloo run bonej2_b200.sharded.thickness_slab with the
oracle-backed stage backend; the assembled result must equal the oracle on the whole volume (maps
bitwise, statistics to 1e-12).  Exercises the z<->y transposes, the r_max halo gather with clipping,
the range-restricted paint, the no-background / no-foreground shortcuts, the statistics combination and the
per-range map checksums bench.py prints."""
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
    from tests.sharded_cpu_backend import OracleStages
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    vol = np.load(vol_path)
    d, h, w = vol.shape
    zs = sharded.split_ranges(d, world)
    slab = torch.from_numpy(vol[zs[rank][0]:zs[rank][1]].copy())
    comm = sharded.Comm()
    timer = sharded.PhaseTimer("cpu")
    out, stats = sharded.thickness_slab(OracleStages(), comm, slab, (w, h, d), inverse, mask, 0.5, timer)
    np.save(os.path.join(out_dir, f"map_{rank}.npy"), out.numpy())
    hashes = sharded.map_hashes(out, zs[rank][0], (w, h, d), parts=world)
    np.save(os.path.join(out_dir, f"hash_{rank}.npy"), np.array([hashes.get(rank, 0)], np.uint64))
    if rank == 0:
        np.save(os.path.join(out_dir, "stats.npy"), np.array([stats["count"], stats["mean"], stats["sd"], stats["min"], stats["max"]]))
        np.save(os.path.join(out_dir, "phases.npy"), np.array(sorted(timer.collect())))
    dist.destroy_process_group()


def _run(vol, inverse, mask, tmp_path, world=2):
    vol_path = os.path.join(tmp_path, "vol.npy")
    np.save(vol_path, vol)
    port = _free_port()
    mp.spawn(_worker, args=(world, port, vol_path, str(tmp_path), inverse, mask), nprocs=world, join=True)
    maps = [np.load(os.path.join(tmp_path, f"map_{r}.npy")) for r in range(world)]
    return np.concatenate(maps, axis=0), np.load(os.path.join(tmp_path, "stats.npy"))


def _check(vol, inverse, mask, tmp_path, world=2):
    from bonej2_b200 import sharded
    from oracle import lt_oracle
    got, st = _run(vol, inverse, mask, tmp_path, world)
    ref = lt_oracle.process_image(vol, inverse=inverse, mask=mask, pixel_width=0.5)
    assert np.array_equal(np.isnan(got), np.isnan(ref["map"]))
    ok = ~np.isnan(got)
    assert np.array_equal(got[ok], ref["map"][ok])
    assert int(st[0]) == ref["stats"].count
    if ref["stats"].count:
        assert math.isclose(st[1], ref["stats"].mean, rel_tol=1e-12)
        assert math.isclose(st[2], ref["stats"].sd, rel_tol=1e-9, abs_tol=1e-9)
        assert st[4] == ref["stats"].max
    # the per-range checksums of the ranks equal those of the whole map computed in one piece (what bench.py compares)
    d, h, w = vol.shape
    whole = sharded.map_hashes(torch.from_numpy(ref["map"].copy()), 0, (w, h, d), parts=world)
    for r in range(world):
        if r in whole:
            assert int(np.load(os.path.join(tmp_path, f"hash_{r}.npy"))[0]) == whole[r]


@pytest.mark.parametrize("inverse", [False, True])
def test_two_ranks_blobs(tmp_path, inverse):
    from tests import shapes
    _check(shapes.random_blobs((17, 14, 19), 3), inverse, True, str(tmp_path))
    phases = [str(x) for x in np.load(os.path.join(str(tmp_path), "phases.npy"))]
    assert {"edt_xy", "transpose", "edt_z", "halo", "ridge", "paint", "finish"} <= set(phases)


def test_four_ranks_thin_slabs(tmp_path):
    """slabs thinner than the largest ball: the halo of a slab reaches across several neighbours"""
    vol = np.zeros((16, 14, 14), np.uint8)
    vol[1:15, 1:13, 1:13] = 255                                         # r_max = 6 > slab thickness 4
    vol[7, 6, 6] = 0
    _check(vol, False, True, str(tmp_path), world=4)


def test_empty_slabs(tmp_path):
    """more ranks than planes: ranks with no plane take part in the collectives only"""
    from tests import shapes
    _check(shapes.random_blobs((3, 9, 8), 2), True, False, str(tmp_path), world=4)


def test_two_ranks_big_radius_halo_clipped(tmp_path):
    """r_max larger than a slab: the halo gather reaches across the whole volume"""
    vol = np.zeros((12, 10, 10), np.uint8)
    vol[:, 1:9, 1:9] = 255
    vol[0] = 0
    _check(vol, False, False, str(tmp_path))


def test_two_ranks_no_background(tmp_path):
    _check(np.full((4, 3, 5), 255, np.uint8), False, False, str(tmp_path))


def test_two_ranks_no_foreground(tmp_path):
    _check(np.zeros((4, 3, 5), np.uint8), False, True, str(tmp_path))


def test_three_ranks_uneven(tmp_path):
    from tests import shapes
    _check(shapes.random_blobs((11, 7, 9), 5), True, True, str(tmp_path), world=3)


def test_split_ranges():
    from bonej2_b200.sharded import split_ranges
    assert split_ranges(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert split_ranges(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]


def _failing_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    from bonej2_b200 import sharded
    from tests.sharded_cpu_backend import OracleStages
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)

    class Stages(OracleStages):
        def paint_range(self, ridge_ext, z_lo, z_hi, max_rsq):
            if rank == 1:
                raise MemoryError("no room for the painter's items on this rank")
            return super().paint_range(ridge_ext, z_lo, z_hi, max_rsq)

    from tests import shapes
    vol = shapes.random_blobs((12, 10, 14), 3)
    zs = sharded.split_ranges(12, world)
    slab = torch.from_numpy(vol[zs[rank][0]:zs[rank][1]].copy())
    try:
        sharded.thickness_slab(Stages(), sharded.Comm(), slab, (14, 10, 12), True, True, 1.0)
        msg = "no error"
    except Exception as e:                                           # noqa: BLE001
        msg = f"{type(e).__name__}: {e}"
    with open(os.path.join(out_dir, f"err_{rank}.txt"), "w") as f:
        f.write(msg)
    dist.destroy_process_group()


def test_failure_on_one_rank_is_raised_on_every_rank(tmp_path):
    """a rank that fails in its paint stage must not leave the others waiting in the statistics collective: it raises its
    own error, the others are told that a rank failed"""
    mp.spawn(_failing_worker, args=(3, _free_port(), str(tmp_path)), nprocs=3, join=True)
    errs = [open(os.path.join(tmp_path, f"err_{r}.txt")).read() for r in range(3)]
    assert errs[1].startswith("MemoryError: no room")
    assert errs[0].startswith("RuntimeError: thickness_slab: 1 other rank(s) failed") and errs[2] == errs[0]
