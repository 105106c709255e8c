"""Hyperstack front-end: subspace order and labels as HyperstackUtilsTest.testViewMeta
(Modern/wrapperPlugins/src/test/java/org/bonej/wrapperPlugins/wrapperUtils/HyperstackUtilsTest.java:326-351),
and batched Thickness per subspace against the oracle (BASELINE.json configs[4])."""
import math

import numpy as np
import pytest

from bonej2_b200 import hyperstack as hs
from bonej2_b200.wrapper import CHANNEL, TIME, UNKNOWN, X, Y, Z
from tests import shapes


def test_view_meta_order_and_labels():
    data = np.zeros((2, 2, 2, 2, 2, 2), np.uint8)                 # numpy order: T, C, W, Z, Y, X
    axes = [X, Y, Z, UNKNOWN, CHANNEL, TIME]
    names = dict(hs.AXIS_NAMES)
    names[UNKNOWN] = "W"
    subs = list(hs.split_3d_subspaces(data, axes, names))
    assert [s.positions for s in subs] == [(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0), (0, 0, 1), (1, 0, 1),
                                           (0, 1, 1), (1, 1, 1)]
    assert [s.label for s in subs] == ["W: 0, Channel: 1, Time: 1", "W: 1, Channel: 1, Time: 1",
                                       "W: 0, Channel: 2, Time: 1", "W: 1, Channel: 2, Time: 1",
                                       "W: 0, Channel: 1, Time: 2", "W: 1, Channel: 1, Time: 2",
                                       "W: 0, Channel: 2, Time: 2", "W: 1, Channel: 2, Time: 2"]
    assert all(s.data.shape == (2, 2, 2) for s in subs)


def test_subspace_content_and_interleaved_axes():
    # ImageJ hyperstack order X, Y, C, Z, T
    c, zdim, t, h, w = 3, 4, 2, 5, 6
    data = np.arange(t * zdim * c * h * w, dtype=np.int64).reshape(t, zdim, c, h, w)
    subs = list(hs.split_3d_subspaces(data, [X, Y, CHANNEL, Z, TIME]))
    assert len(subs) == c * t
    assert subs[0].label == "Channel: 1, Time: 1" and subs[1].label == "Channel: 2, Time: 1"
    assert subs[3].label == "Channel: 1, Time: 2"
    assert np.array_equal(subs[4].data, data[1, :, 1])            # C = 2, T = 2 -> [z][y][x]


def test_same_type_axes_get_subscripts():
    data = np.zeros((2, 2, 3, 3, 3), np.uint8)
    subs = list(hs.split_3d_subspaces(data, [X, Y, Z, CHANNEL, CHANNEL]))
    assert subs[1].label == "Channel: 2, Channel(2): 1"


@pytest.mark.gpu
def test_batched_thickness_matches_oracle(oracle):
    vols = [shapes.random_blobs((12, 14, 16), s) for s in (1, 2, 3)] + [np.zeros((12, 14, 16), np.uint8)]
    data = np.stack(vols).reshape(2, 2, 12, 14, 16)               # T, C, Z, Y, X
    rows = hs.thickness_hyperstack(data, [X, Y, Z, CHANNEL, TIME], name="stack", map_choice="Both", unit="mm")
    assert [r[0] for r in rows] == ["stack Channel: 1, Time: 1", "stack Channel: 2, Time: 1",
                                    "stack Channel: 1, Time: 2", "stack Channel: 2, Time: 2"]
    for (label, vals), vol in zip(rows, vols):
        for prefix, inverse in (("Tb.Th", False), ("Tb.Sp", True)):
            st = oracle.process_image(vol, inverse=inverse, mask=True)["stats"]
            got = vals[f"{prefix} Mean (mm)"]
            if st.count == 0:
                assert math.isnan(got)
            else:
                assert math.isclose(got, st.mean, rel_tol=1e-5)
                assert math.isclose(vals[f"{prefix} Max (mm)"], st.max, rel_tol=1e-5)


def _ranks_worker(rank, world, port, path, out_dir):
    import os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch.distributed as dist
    from bonej2_b200 import hyperstack as H
    from oracle import lt_oracle
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    data = np.load(path)
    seen = []

    def run(vol, inverse):                                      # oracle stands in for the CUDA call: the test is about
        seen.append(int(vol.sum()))                             # who runs which subspace and the order of the rows
        st = lt_oracle.process_image(vol, inverse=inverse, mask=True)["stats"]
        return (st.mean, st.sd, st.max) if st.count else (float("nan"),) * 3

    rows = H.thickness_hyperstack_ranks(data, [X, Y, Z, CHANNEL, TIME], name="s", map_choice="Trabecular thickness", run=run)
    np.save(os.path.join(out_dir, f"rows_{rank}.npy"), np.array([[r[1]["Tb.Th Mean (pixel)"] for r in rows]]))
    with open(os.path.join(out_dir, f"labels_{rank}.txt"), "w") as f:
        f.write("\n".join(r[0] for r in rows) + f"\n#{len(seen)}")
    dist.destroy_process_group()


def test_ranks_split_subspaces_and_agree(tmp_path, oracle):
    """world_size 2 over gloo: each rank runs every other subspace, both end with all rows in subspace order"""
    import socket
    import torch.multiprocessing as mp
    vols = [shapes.random_blobs((8, 9, 10), s) for s in (1, 2, 3, 4, 5)] + [np.zeros((8, 9, 10), np.uint8)]
    data = np.stack(vols).reshape(2, 3, 8, 9, 10)               # T, C, Z, Y, X -> 6 subspaces
    path = str(tmp_path / "hs.npy")
    np.save(path, data)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_ranks_worker, args=(2, port, path, str(tmp_path)), nprocs=2, join=True)
    want = []
    for v in vols:
        st = oracle.process_image(v, inverse=False, mask=True)["stats"]
        want.append(st.mean if st.count else float("nan"))
    for rank in (0, 1):
        got = np.load(str(tmp_path / f"rows_{rank}.npy"))[0]
        assert np.allclose(got, want, rtol=0, atol=0, equal_nan=True)
        lines = open(str(tmp_path / f"labels_{rank}.txt")).read().split("\n")
        assert lines[:3] == ["s Channel: 1, Time: 1", "s Channel: 2, Time: 1", "s Channel: 3, Time: 1"]
        assert lines[-1] == "#3"                                 # three of the six subspaces ran here


@pytest.mark.gpu
def test_ranks_single_process_cuda(oracle):
    """without a process group the rank form is the plain batch on this GPU"""
    vols = [shapes.random_blobs((12, 14, 16), s) for s in (4, 5)]
    rows = hs.thickness_hyperstack_ranks(np.stack(vols), [X, Y, Z, TIME], name="t", map_choice="Both")
    assert [r[0] for r in rows] == ["t Time: 1", "t Time: 2"]
    for (label, vals), vol in zip(rows, vols):
        st = oracle.process_image(vol, inverse=True, mask=True)["stats"]
        assert math.isclose(vals["Tb.Sp Mean (pixel)"], st.mean, rel_tol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("map_choice", ["Both", "Trabecular separation"])
def test_devices_work_concurrently_and_agree(map_choice):
    """three contexts (here on the GPUs of the box, sharing one if there is only one), one host thread each: rows and
    maps are those of the single-context batch, in subspace order"""
    import torch
    vols = [shapes.random_blobs((16, 18, 20), s) for s in (1, 2, 3, 4, 5)] + [np.zeros((16, 18, 20), np.uint8),
                                                                           np.full((16, 18, 20), 255, np.uint8)]
    data = np.stack(vols)                                          # T = 7, Z, Y, X
    ndev = torch.cuda.device_count()
    one, maps1 = hs.thickness_hyperstack(data, [X, Y, Z, TIME], name="h", map_choice=map_choice, want_maps=True)
    many, maps3 = hs.thickness_hyperstack(data, [X, Y, Z, TIME], name="h", map_choice=map_choice, want_maps=True,
                                          devices=[i % ndev for i in range(3)])
    assert [r[0] for r in one] == [r[0] for r in many] == [f"h Time: {t}" for t in range(1, 8)]
    for (_, a), (_, b), ma, mb in zip(one, many, maps1, maps3):
        assert a.keys() == b.keys()
        for k in a:
            assert (math.isnan(a[k]) and math.isnan(b[k])) or a[k] == b[k], k
        for prefix in ma:
            assert np.array_equal(ma[prefix].view(np.uint32), mb[prefix].view(np.uint32))


@pytest.mark.gpu
def test_failure_on_one_device_is_raised():
    from bonej2_b200._native import ThicknessError
    vols = [shapes.random_blobs((8, 9, 10), 1), np.full((8, 9, 10), 7, np.uint8)]      # the second is not binary
    with pytest.raises(ThicknessError):
        hs.thickness_hyperstack(np.stack(vols), [X, Y, Z, TIME], devices=[0, 0])
