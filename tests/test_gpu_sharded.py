"""bonej2_b200.sharded with the CUDA backend: one rank on one GPU must reproduce the oracle (and so the
single-call path); with >= 2 GPUs two NCCL ranks must reproduce it too, bitwise."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

from tests import shapes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cmp(got, st, ref):
    assert np.array_equal(np.isnan(got), np.isnan(ref["map"]))
    ok = ~np.isnan(got)
    assert np.allclose(got[ok], ref["map"][ok], rtol=1e-5, atol=0)
    assert st["count"] == ref["stats"].count
    if st["count"]:
        assert abs(st["mean"] - ref["stats"].mean) <= 1e-5 * abs(ref["stats"].mean)
        assert st["max"] == pytest.approx(ref["stats"].max, rel=1e-6)


@pytest.mark.parametrize("use_peer", [False, True], ids=["plain", "peer_buffers"])
@pytest.mark.parametrize("inverse", [False, True])
def test_single_rank_matches_oracle(ctx, oracle, inverse, use_peer):
    """one rank: the slab is the volume.  With use_peer the layout changes go through the exchange buffers and the
    strided device copies of the multi-rank path (bjt_dev_alloc, bjt_dev_copy2d), pulling from itself."""
    from bonej2_b200 import sharded
    vol = shapes.random_blobs((40, 36, 44), 9)
    dev = torch.device("cuda", 0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    slab = torch.from_numpy(vol).to(dev)
    d, h, w = vol.shape
    comm = sharded.Comm()
    peer = sharded.PeerExchange(ctx, comm, (w, h, d), dev) if use_peer else None
    out, st = sharded.thickness_slab(sharded.CudaStages(ctx, dev), comm, slab, (w, h, d), inverse, True, 0.7, None, peer)
    _cmp(out.cpu().numpy(), st, oracle.process_image(vol, inverse=inverse, mask=True, pixel_width=0.7))


def test_ipc_handle_round_trip_in_process(ctx):
    """bjt_ipc_export gives a 64-byte handle of a library buffer; importing it in the exporting process is refused by
    CUDA (a handle is for another process) and reported as an error, never ignored"""
    from bonej2_b200 import _native
    p = ctx.dev_alloc("ipc_test", 4096)
    h = ctx.ipc_export(p)
    assert len(h) == 64 and any(h)
    assert ctx.dev_alloc("ipc_test", 1024) == p                      # named buffers persist and only grow
    with pytest.raises(_native.ThicknessError):
        ctx.ipc_import(h)


def _worker(rank, world, port, vol_path, out_dir, inverse, use_peer):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from bonej2_b200 import _native, sharded
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    vol = np.load(vol_path)
    d, h, w = vol.shape
    zs = sharded.split_ranges(d, world)
    dev = torch.device("cuda", rank)
    c = _native.Context(rank)
    c.set_stream(torch.cuda.current_stream().cuda_stream)
    slab = torch.from_numpy(vol[zs[rank][0]:zs[rank][1]].copy()).to(dev)
    comm = sharded.Comm()
    peer = sharded.PeerExchange(c, comm, (w, h, d), dev) if use_peer else None
    stages = sharded.CudaStages(c, dev)
    if use_peer:    # the other run first: the exchange buffers are reused from run to run, as in a "Both" call
        sharded.thickness_slab(stages, comm, slab, (w, h, d), not inverse, True, 1.0, None, peer)
    out, st = sharded.thickness_slab(stages, comm, slab, (w, h, d), inverse, True, 1.0, None, peer)
    np.save(os.path.join(out_dir, f"map_{rank}.npy"), out.cpu().numpy())
    if peer is not None:
        torch.cuda.synchronize(); comm.barrier()
        peer.close()
    if rank == 0:
        np.save(os.path.join(out_dir, "stats.npy"), np.array([st["count"], st["mean"], st["sd"], st["min"], st["max"]]))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("use_peer", [False, True], ids=["nccl_p2p", "peer_pulls"])
@pytest.mark.parametrize("inverse", [False, True])
def test_two_ranks_nccl_bitwise(oracle, tmp_path, inverse, use_peer):
    """two ranks, two GPUs: layout changes by NCCL send / recv, and by pulls out of the peer's memory (CUDA IPC)"""
    import torch.multiprocessing as mp
    from bonej2_b200 import _native
    from bonej2_b200.phantom import phantom_shapes
    n = 96
    vol = oracle.phantom_raster(phantom_shapes(n, n, n, 2), n, n, n)
    vol_path = os.path.join(tmp_path, "vol.npy")
    np.save(vol_path, vol)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, vol_path, str(tmp_path), inverse, use_peer), nprocs=2, join=True)
    got = np.concatenate([np.load(os.path.join(tmp_path, f"map_{r}.npy")) for r in range(2)], axis=0)
    with _native.Context(0) as c:
        one, st1, _ = c.thickness(vol, inverse=inverse, mask=True)
    assert np.array_equal(np.nan_to_num(got, nan=-1.0), np.nan_to_num(one, nan=-1.0)), "2-GPU map != 1-GPU map"
    st = np.load(os.path.join(tmp_path, "stats.npy"))
    assert int(st[0]) == st1.count and abs(st[1] - st1.mean) <= 1e-9 * abs(st1.mean)


@pytest.mark.parametrize("shape,cuts,seed", [((48, 40, 44), [0, 13, 30, 48], 4), ((20, 24, 200), [0, 4, 9, 14, 20], 6)])
@pytest.mark.parametrize("inverse", [False, True])
def test_paint_range_slabs_one_gpu(ctx, oracle, shape, cuts, seed, inverse):
    """bjt_dev_paint_range on z-slabs of one volume: every slab is painted from the ridge of slab + halo (r_max planes
    either side, as bonej2_b200.sharded fetches them); the slabs' painted r^2 must equal the oracle's on the whole
    volume, bitwise.  The second case has slabs thinner than the largest ball."""
    import math
    vol = shapes.random_blobs(shape, seed)
    dist_, sq = oracle.edt(vol, inverse=inverse)
    rg = oracle.ridge(dist_)
    _, want = oracle.paint(rg)
    ridge = np.where(rg > 0, sq, 0).astype(np.int32)
    gmax = int(ridge.max())
    H = math.isqrt(gmax)
    d, h, w = shape
    dev = torch.device("cuda", 0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    got = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        e_lo, e_hi = max(0, a - H), min(d, b + H)
        ext = torch.from_numpy(ridge[e_lo:e_hi].copy()).to(dev)
        out = torch.empty((b - a, h, w), dtype=torch.int32, device=dev)
        ctx.dev_paint_range(ext.data_ptr(), w, h, e_hi - e_lo, a - e_lo, b - e_lo, gmax, out.data_ptr())
        got.append(out.cpu().numpy())
    assert np.array_equal(np.concatenate(got).astype(np.uint32), want)


class _Stepwise:
    """thin driver of the stepwise painter entry points (bjt_dev_paint_open .. _close) for the test below"""
    KCAP, MAX_SOURCES = 32, 4

    def __init__(self, ctx, device):
        self.ctx, self.device = ctx, device

    def open(self, ridge_own, max_rsq):
        dz, h, w = ridge_own.shape
        self.ridge, self.out = ridge_own, torch.empty((dz, h, w), dtype=torch.int32, device=self.device)
        return self.ctx.dev_paint_open(ridge_own.data_ptr(), w, h, dz, max_rsq)

    def cols(self, xs):
        return self.ctx.dev_paint_xslab_width(xs) * self.ridge.shape[1]

    def buffer(self, ncol):
        return torch.empty(ncol * (2 * self.KCAP + 1), dtype=torch.int32, device=self.device)

    def split(self, buf, ncol):
        return buf.data_ptr() + 8 * ncol * self.KCAP, buf.data_ptr()              # (count pointer, items pointer)

    def export(self, xs, side, gap, nplanes, buf, ncol):
        cnt, items = self.split(buf, ncol)
        self.ctx.dev_paint_export(xs, side, gap, nplanes, cnt, items, self.KCAP)

    def sweep(self, xs, lo, hi, ncol):
        self.ctx.dev_paint_sweep(xs, [self.split(b, ncol) for b in lo], [self.split(b, ncol) for b in hi], self.KCAP,
                                 self.out.data_ptr())


@pytest.mark.parametrize("shape,cuts,seed,xcols", [((48, 40, 44), [0, 13, 30, 48], 4, 0),       # three slabs
                                                    ((20, 24, 200), [0, 4, 9, 14, 20], 6, 64)])  # four column blocks, thin z-slabs
@pytest.mark.parametrize("inverse,wide", [(False, False), (True, False), (True, True)])
def test_stepwise_painter_slabs_one_gpu(oracle, shape, cuts, seed, xcols, inverse, wide):
    """The sweep painter in steps (bjt_dev_paint_open .. _close, the large-radius alternative to bjt_dev_paint_range) on
    z-slabs of one volume, all on this GPU, one context per slab, boundary staircases handed over in device memory:
    the slabs' painted r^2 must equal the oracle's on the whole volume, bitwise."""
    import math
    from bonej2_b200 import _native
    vol = shapes.random_blobs(shape, seed)
    dist_, sq = oracle.edt(vol, inverse=inverse)
    rg = oracle.ridge(dist_)
    _, want = oracle.paint(rg)
    ridge = np.where(rg > 0, sq, 0).astype(np.int32)
    gmax = int(ridge.max())
    rmax = math.isqrt(gmax)
    tag_bound = max(gmax, 70000) if wide else gmax              # a looser bound is legal and selects the u64-item kernels
    d, h, w = shape
    dev = torch.device("cuda", 0)
    zs = list(zip(cuts[:-1], cuts[1:]))
    ctxs = [_native.Context(0) for _ in zs]
    _native.Context.debug_set_xslab_cols(xcols)
    try:
        st = []
        for c, (a, b) in zip(ctxs, zs):
            c.set_stream(torch.cuda.current_stream().cuda_stream)
            s = _Stepwise(c, dev)
            s.nxs = s.open(torch.from_numpy(ridge[a:b].copy()).to(dev), tag_bound)
            st.append(s)
        nxs = st[0].nxs
        assert nxs == ((w + xcols - 1) // xcols if xcols else 1)

        def gap(p, q):
            g = zs[q][0] - (zs[p][1] - 1) if q > p else zs[p][0] - (zs[q][1] - 1)
            return g if g <= rmax else 0

        for xs in range(nxs):
            ncol = st[0].cols(xs)
            for s in st:
                s.ctx.dev_paint_lists(xs)
            box = {}
            for p, s in enumerate(st):
                for q in range(len(st)):
                    if q != p and gap(p, q):
                        box[(p, q)] = s.buffer(ncol)
                        s.export(xs, +1 if q > p else -1, gap(p, q), rmax + 1, box[(p, q)], ncol)
            for q, s in enumerate(st):
                lo = [box[(p, q)] for p in range(q) if (p, q) in box]
                hi = [box[(p, q)] for p in range(q + 1, len(st)) if (p, q) in box]
                assert len(lo) <= s.MAX_SOURCES and len(hi) <= s.MAX_SOURCES
                s.sweep(xs, lo, hi, ncol)
        got = []
        for s in st:
            assert s.ctx.dev_paint_close() == 0
            got.append(s.out.cpu().numpy())
        assert np.array_equal(np.concatenate(got).astype(np.uint32), want)
    finally:
        _native.Context.debug_set_xslab_cols(0)
        for c in ctxs:
            c.close()
