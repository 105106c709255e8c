"""z-slab sharded Thickness: one process per GPU, torch.distributed for the plumbing.

The volume [d][h][w] is cut into contiguous z-slabs, one per rank.  Per pipeline run:

  EDT x, y          slab-local kernels
  EDT z             all-to-all transpose z-slabs -> y-slabs, z pass on whole z lines, transpose back
  r_max             all-reduce of the largest squared distance: no ball reaches further than H = floor(sqrt(max))
  halo              ONE exchange per run: every rank fetches the squared-EDT planes within H + 3 of its slab (the
                    "max-radius halo exchange" of BASELINE.json's north_star)
                    On GPUs with even splits both layout changes are PULLS out of the peers' HBM (PeerExchange: the
                    y-pass and z-pass results live in buffers every other rank has mapped through CUDA IPC; a rank
                    copies its share with one strided device-to-device copy per peer over NVLink, and the transpose
                    back fetches slab + halo planes in the same copy, straight from the y-slab layout): no pack or
                    unpack kernels, no staging, no separate halo exchange.  Otherwise (CPU tests over gloo, ragged
                    splits) the same data moves with point-to-point sends.
  ridge             on slab +- (H + 2) planes (the outermost fetched plane only serves as a neighbour)
  paint             planes slab +- 2 only (bjt_dev_paint_range): the painter's fronts run along z and read the ridge
                    within H of a plane, its discs are plane-local, so nothing else crosses a slab face
  cleanup .. stats  on slab +- 2 painted planes, output and statistics restricted to the slab
                    (bjt_dev_finish_range); (n, sum, sum^2, min, max) are combined across ranks before
                    mean and sd are formed, as ij.process.StackStatistics does on the whole stack

The result is bitwise the single-GPU map (integer stages are order-independent, the cleanup sees the
same neighbours).  The compute backend is pluggable so the exchange logic can be tested on CPU with
world_size 2 over gloo (tests/test_sharded_cpu.py injects an oracle-backed backend); the product
backend below is CUDA only -- there is no CPU fallback.
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass

import torch
import torch.distributed as dist


def split_ranges(n: int, parts: int):
    """contiguous near-equal ranges [(lo, hi)] of range(n)"""
    base, rem = divmod(n, parts)
    out, lo = [], 0
    for i in range(parts):
        hi = lo + base + (1 if i < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


@dataclass
class PartialStats:
    count: int
    sum: float
    sum2: float
    min: float
    max: float

    def finalize(self):
        """StackStatistics: mean = sum/n, sd = sqrt((n*sum2 - sum^2)/n/(n-1)) clamped at 0"""
        n = float(self.count)
        mean = self.sum / n if self.count else float("nan")
        sd = 0.0
        if self.count > 0:
            v = (n * self.sum2 - self.sum * self.sum) / n
            sd = math.sqrt(v / (n - 1.0)) if v > 0.0 else 0.0
        return dict(count=self.count, mean=mean, sd=sd, min=self.min, max=self.max)


class Comm:
    """the few collectives the path needs, on whatever backend the process group has"""

    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.size = dist.get_world_size(group) if dist.is_initialized() else 1

    def exchange(self, sends, recvs):
        """sends / recvs: {peer: contiguous tensor}.  Point-to-point, works on nccl and gloo alike."""
        if self.size == 1:
            return
        ops = []
        for peer, t in sorted(recvs.items()):
            ops.append(dist.P2POp(dist.irecv, t, peer, self.group))
        for peer, t in sorted(sends.items()):
            ops.append(dist.P2POp(dist.isend, t, peer, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def allreduce_max_int(self, v: int, device) -> int:
        if self.size == 1:
            return v
        t = torch.tensor([v], dtype=torch.int64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return int(t.item())

    def combine_stats(self, st: PartialStats, device, failed=False):
        """(statistics over all ranks, number of ranks that report a failure)"""
        if self.size == 1:
            return st, int(failed)
        t = torch.tensor([float(st.count), st.sum, st.sum2, st.min, st.max, 1.0 if failed else 0.0], dtype=torch.float64,
                         device=device)
        all_t = [torch.empty_like(t) for _ in range(self.size)]
        dist.all_gather(all_t, t, group=self.group)
        rows = [x.cpu().tolist() for x in all_t]                      # rank order: the same sum on every rank
        return (PartialStats(int(sum(r[0] for r in rows)), sum(r[1] for r in rows), sum(r[2] for r in rows),
                             min(r[3] for r in rows), max(r[4] for r in rows)), int(sum(r[5] for r in rows)))

    def barrier(self):
        if self.size > 1:
            dist.barrier(group=self.group)

    def stream_barrier(self, device):
        """a collective in stream order: when it has run on this rank's stream, every rank's work queued before its own
        call has finished (what a pull out of a peer's buffer waits for)"""
        if self.size > 1:
            t = torch.zeros(1, dtype=torch.int32, device=device)
            dist.all_reduce(t, group=self.group)


class DeviceArray:
    """a device buffer of the library (bjt_dev_alloc) or of a peer (bjt_ipc_import) as something torch can view"""

    def __init__(self, ptr, shape, typestr):
        self.ptr = ptr
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}


class PeerExchange:
    """The two buffers of this rank that the other ranks read -- gy [dz][h][w] (y-pass result, z-slab layout) and
    sqt [d][hy][w] (z-pass result, y-slab layout) -- and the mappings of every peer's pair (CUDA IPC: the ranks are
    processes).  Needs d and h divisible by the number of ranks (every rank's buffers then have the same shape)."""

    def __init__(self, ctx, comm, dims, device):
        w, h, d = dims
        G, r = comm.size, comm.rank
        if d % G or h % G:
            raise ValueError("PeerExchange needs d and h divisible by the number of ranks")
        self.ctx, self.dims, self.dz, self.hy = ctx, dims, d // G, h // G
        self.gy_ptr = ctx.dev_alloc("peer_gy", self.dz * h * w * 4)
        self.sqt_ptr = ctx.dev_alloc("peer_sqt", d * self.hy * w * 4)
        self.gy = torch.as_tensor(DeviceArray(self.gy_ptr, (self.dz, h, w), "<i4"), device=device)
        self.sqt = torch.as_tensor(DeviceArray(self.sqt_ptr, (d, self.hy, w), "<i4"), device=device)
        mine = (ctx.ipc_export(self.gy_ptr), ctx.ipc_export(self.sqt_ptr))
        everyone = [mine]
        if G > 1:
            everyone = [None] * G
            dist.all_gather_object(everyone, mine, group=comm.group)
        self.peer_gy, self.peer_sqt, self._opened = [0] * G, [0] * G, []
        for p in range(G):
            if p == r:
                self.peer_gy[p], self.peer_sqt[p] = self.gy_ptr, self.sqt_ptr
            else:
                self.peer_gy[p] = ctx.ipc_import(everyone[p][0])
                self.peer_sqt[p] = ctx.ipc_import(everyone[p][1])
                self._opened += [self.peer_gy[p], self.peer_sqt[p]]

    def close(self):
        for ptr in self._opened:
            self.ctx.ipc_close(ptr)
        self._opened = []

    def pull_y_slab(self, comm, out):
        """out [d][hy][w] <- my y range of every rank's gy (the all-to-all transpose z-slabs -> y-slabs)"""
        w, h, d = self.dims
        G, r, dz, hy = comm.size, comm.rank, self.dz, self.hy
        row = hy * w * 4
        order = [(r + i) % G for i in range(G)]                        # every rank starts with a different peer
        self.ctx.dev_copy2d_batch([(out.data_ptr() + p * dz * row, row, self.peer_gy[p] + r * row, h * w * 4, row, dz) for p in order])

    def pull_planes(self, comm, out, lo, hi):
        """out [hi - lo][h][w] <- planes [lo, hi) of the squared EDT, gathered from every rank's y-slab result (the
        transpose back and the halo in one copy per peer)"""
        w, h, d = self.dims
        G, r, hy = comm.size, comm.rank, self.hy
        row = hy * w * 4
        order = [(r + i) % G for i in range(G)]
        self.ctx.dev_copy2d_batch([(out.data_ptr() + p * row, h * w * 4, self.peer_sqt[p] + lo * row, row, row, hi - lo) for p in order])


# ------------------------------------------------------------------------------------------------------
# layout changes

def transpose_z_to_y(comm: Comm, t, zs, ys):
    """[dz_r][h][w] on every rank r  ->  [d][hy_r][w]: rank r ends up with all z of its y range"""
    r = comm.rank
    d_total = zs[-1][1]
    w = t.shape[2]
    out = t.new_empty((d_total, ys[r][1] - ys[r][0], w))
    sends, recvs, bufs = {}, {}, {}
    for p in range(comm.size):
        chunk = t[:, ys[p][0]:ys[p][1], :]
        if p == r:
            out[zs[r][0]:zs[r][1]] = chunk
        else:
            if chunk.numel():
                sends[p] = chunk.contiguous()
            if (zs[p][1] - zs[p][0]) * (ys[r][1] - ys[r][0]) > 0:
                recvs[p] = out[zs[p][0]:zs[p][1]]                   # contiguous: z is the slowest axis
    comm.exchange(sends, recvs)
    return out


def transpose_y_to_z(comm: Comm, t, zs, ys, h):
    """inverse of transpose_z_to_y: [d][hy_r][w] -> [dz_r][h][w]"""
    r = comm.rank
    w = t.shape[2]
    out = t.new_empty((zs[r][1] - zs[r][0], h, w))
    sends, recvs, tmp = {}, {}, {}
    for p in range(comm.size):
        chunk = t[zs[p][0]:zs[p][1]]                                  # what rank p's slab needs of my y range
        if p == r:
            out[:, ys[r][0]:ys[r][1], :] = chunk
        else:
            if chunk.numel():
                sends[p] = chunk.contiguous()
            n = (zs[r][1] - zs[r][0]) * (ys[p][1] - ys[p][0])
            if n > 0:
                tmp[p] = t.new_empty((zs[r][1] - zs[r][0], ys[p][1] - ys[p][0], w))
                recvs[p] = tmp[p]
    comm.exchange(sends, recvs)
    for p, buf in tmp.items():
        out[:, ys[p][0]:ys[p][1], :] = buf
    return out


def gather_planes(comm: Comm, slab, zs, needs):
    """needs[r] = (lo, hi): the global plane range rank r wants.  Returns this rank's [hi-lo][h][w]."""
    r = comm.rank
    lo, hi = needs[r]
    out = slab.new_empty((hi - lo,) + tuple(slab.shape[1:]))
    sends, recvs = {}, {}
    for p in range(comm.size):
        # what I hold of p's wish
        a, b = max(needs[p][0], zs[r][0]), min(needs[p][1], zs[r][1])
        if p == r:
            if b > a:
                out[a - lo:b - lo] = slab[a - zs[r][0]:b - zs[r][0]]
            continue
        if b > a:
            sends[p] = slab[a - zs[r][0]:b - zs[r][0]].contiguous()
        # what p holds of my wish
        a2, b2 = max(lo, zs[p][0]), min(hi, zs[p][1])
        if b2 > a2:
            recvs[p] = out[a2 - lo:b2 - lo]
    comm.exchange(sends, recvs)
    return out


# ------------------------------------------------------------------------------------------------------
# the product backend: CUDA kernels through the C-ABI on torch device tensors

class CudaStages:
    def __init__(self, ctx, device):
        self.ctx, self.device = ctx, device
        self._bufs = {}

    def scratch(self, name, shape, dtype):
        """A per-stage work buffer that is kept between runs and only ever grows.  The sizes of the halo buffers change
        from run to run (they follow the largest radius), and every fresh device allocation of a process that has peer
        mappings open is mapped into all peers: at 2048^3 on 8 GPUs the ridge phase took 62 ms with torch.empty per
        run against 18 ms before the peer buffers existed."""
        n = 1
        for k in shape:
            n *= int(k)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        t = self._bufs.get(name)
        if t is None or t.numel() < nbytes:
            t = None
            self._bufs[name] = None
            t = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=self.device)
            self._bufs[name] = t
        return t[:nbytes].view(dtype).view(tuple(int(k) for k in shape))

    def _i32(self, shape, name=None):
        if name is not None:
            return self.scratch(name, shape, torch.int32)
        return torch.empty(shape, dtype=torch.int32, device=self.device)

    def edt_x(self, slab_u8, inverse, threshold=128):
        dz, h, w = slab_u8.shape
        gx = self.scratch("gx", (dz, h, w), torch.int16)                           # u16 bits
        self.ctx.dev_edt_x(slab_u8.data_ptr(), w, h, dz, inverse, threshold, gx.data_ptr())
        return gx

    def edt_y(self, gx, n_sentinel, out=None):
        dz, h, w = gx.shape
        gy = self._i32((dz, h, w)) if out is None else out
        self.ctx.dev_edt_y(gx.data_ptr(), w, h, dz, n_sentinel, gy.data_ptr())
        return gy

    def edt_z(self, gy_t, n_sentinel, out=None):
        d, hy, w = gy_t.shape
        sq = self._i32((d, hy, w)) if out is None else out
        if hy > 0:
            self.ctx.dev_edt_z(gy_t.data_ptr(), w, hy, d, n_sentinel, sq.data_ptr())
        return sq

    def max_value(self, t):
        return int(t.max().item()) if t.numel() else 0

    def ridge(self, sq_ext):
        de, h, w = sq_ext.shape
        out = self._i32((de, h, w), "ridge_ext")
        self.ctx.dev_ridge(sq_ext.data_ptr(), w, h, de, out.data_ptr())
        return out

    def paint_range(self, ridge_ext, z_lo, z_hi, max_rsq):
        """painted r^2 of planes [z_lo, z_hi) of ridge_ext (balls centred on any plane of ridge_ext count)"""
        de, h, w = ridge_ext.shape
        out = self._i32((z_hi - z_lo, h, w), "rsq_sub")
        self.ctx.dev_paint_range(ridge_ext.data_ptr(), w, h, de, z_lo, z_hi, max_rsq, out.data_ptr())
        return out

    def full(self, shape, value):
        return torch.full(shape, value, dtype=torch.int32, device=self.device)

    def finish(self, rsq_sub, orig_core_u8, z_lo, z_hi, inverse, mask, pixel_width):
        ds, h, w = rsq_sub.shape
        out = torch.empty((z_hi - z_lo, h, w), dtype=torch.float32, device=self.device)
        # the kernel indexes the original with the coordinates of rsq_sub: shift the core pointer
        optr = (orig_core_u8.data_ptr() - z_lo * h * w) if mask else None
        st = self.ctx.dev_finish_range(rsq_sub.data_ptr(), optr, w, h, ds, z_lo, z_hi, inverse, True, pixel_width,
                                       out.data_ptr())
        return out, PartialStats(st.count, st.sum, st.sum2, st.min, st.max)

    def paint_range_into(self, ridge_ext, z_lo, z_hi, max_rsq, out_view):
        """paint_range into planes of an existing buffer (out_view: [z_hi - z_lo][h][w], contiguous)"""
        de, h, w = ridge_ext.shape
        self.ctx.dev_paint_range(ridge_ext.data_ptr(), w, h, de, z_lo, z_hi, max_rsq, out_view.data_ptr())

    def finish_range_into(self, rsq_sub, orig_core_u8, core_lo, z_lo, z_hi, inverse, mask, pixel_width, out_view):
        """finish planes [z_lo, z_hi) of rsq_sub (its planes within 2 of them must be painted) into out_view; core_lo =
        the plane of rsq_sub that is the first plane of orig_core_u8"""
        ds, h, w = rsq_sub.shape
        optr = (orig_core_u8.data_ptr() - core_lo * h * w) if mask else None
        st = self.ctx.dev_finish_range(rsq_sub.data_ptr(), optr, w, h, ds, z_lo, z_hi, inverse, True, pixel_width,
                                       out_view.data_ptr())
        return PartialStats(st.count, st.sum, st.sum2, st.min, st.max)


def add_partials(a: "PartialStats", b: "PartialStats") -> "PartialStats":
    if b.count == 0:
        return a
    if a.count == 0:
        return b
    return PartialStats(a.count + b.count, a.sum + b.sum, a.sum2 + b.sum2, min(a.min, b.min), max(a.max, b.max))


# ------------------------------------------------------------------------------------------------------

def no_result(n: int) -> int:
    return 3 * (n + 1) * (n + 1)


class PhaseTimer:
    """device time per phase from CUDA events on the current stream (no synchronisation inside the timed loop);
    on CPU backends wall-clock time"""

    def __init__(self, device):
        self.cuda = str(device).startswith("cuda")
        self.marks, self.ms = [], {}

    def mark(self, name):
        if self.cuda:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.marks.append((name, ev))
        else:
            self.marks.append((name, time.perf_counter()))

    def collect(self):
        """call after a synchronize: accumulates the intervals between consecutive marks under the later mark's name"""
        for (_, a), (name, b) in zip(self.marks, self.marks[1:]):
            if name != "start":
                self.ms[name] = self.ms.get(name, 0.0) + (a.elapsed_time(b) if self.cuda else 1e3 * (b - a))
        self.marks = []
        return self.ms


def thickness_slab(stages, comm: Comm, slab_u8, dims, inverse, mask=True, pixel_width=1.0, timer=None, peer=None, sink=None):
    """One pipeline run (LocalThicknessWrapper.processImage + StackStatistics) on this rank's z-slab.
    slab_u8: [dz][h][w] planes zs[rank] of the volume dims = (w, h, d).  Returns (map slab f32, stats dict).
    peer: a PeerExchange for these dims -- the layout changes are then pulls out of the peers' memory.
    sink(a, b, view): called, in stream order, as soon as planes [a, b) of the map slab are final (view = those planes
    of the returned tensor): the slab is then painted and finished in chunks of planes so that a caller can send each
    chunk on its way (to the host, say) while the next one is painted."""
    w, h, d = dims
    G, r = comm.size, comm.rank
    zs, ys = split_ranges(d, G), split_ranges(h, G)
    z0, z1 = zs[r]
    n = max(w, h, d)
    n0 = no_result(n)
    mark = timer.mark if timer is not None else (lambda name: None)

    mark("start")
    gx = stages.edt_x(slab_u8, inverse)
    if peer is not None:
        stages.edt_y(gx, n, out=peer.gy)
        del gx
        mark("edt_xy")
        comm.stream_barrier(slab_u8.device)                            # every rank's y pass has finished
        gy_t = stages.scratch("gy_t", (d, ys[r][1] - ys[r][0], w), torch.int32)
        peer.pull_y_slab(comm, gy_t)
        mark("transpose")
        stages.edt_z(gy_t, n, out=peer.sqt)
        del gy_t
        mark("edt_z")
        # the collective also orders the pulls: after it every rank's z pass has finished (sqt may be read) and every
        # pull out of gy has completed (the next run may overwrite it)
        gmax = comm.allreduce_max_int(stages.max_value(peer.sqt), slab_u8.device)
        mark("rmax")
        sq = None
    else:
        gy = stages.edt_y(gx, n)
        del gx
        mark("edt_xy")
        gy_t = transpose_z_to_y(comm, gy, zs, ys)
        del gy
        mark("transpose")
        sq_t = stages.edt_z(gy_t, n)
        del gy_t
        mark("edt_z")
        sq = transpose_y_to_z(comm, sq_t, zs, ys, h)
        del sq_t
        mark("transpose")
        gmax = comm.allreduce_max_int(stages.max_value(sq), slab_u8.device)
        mark("rmax")

    # An error on one rank in the ridge / paint / finish stages (out of memory in the painter, a library failure) must
    # not leave the others waiting in the next collective: it is caught, travels with the statistics, and every rank
    # raises.  (With a PeerExchange nothing in this section waits for another rank; without one the halo exchange does,
    # so there the guarantee covers the stages behind it.)
    failure = None
    streamed_out = None
    try:
        # planes of the painted map the cleanup of this slab reads: slab +- 2, clipped to the volume
        s_lo, s_hi = (max(0, z0 - 2), min(d, z1 + 2)) if z1 > z0 else (z0, z0)
        if gmax == 0:
            rsq_sub = stages.full((s_hi - s_lo, h, w), 0)
        elif gmax >= n0:
            rsq_sub = stages.full((s_hi - s_lo, h, w), n0)                 # no background anywhere (sentinel)
        else:
            H = math.isqrt(gmax)                                           # no ball reaches further
            # painting [s_lo, s_hi) reads the ridge of [s_lo - H, s_hi + H); the ridge of a plane needs the squared
            # EDT one plane further out
            needs = [((max(0, a - 2 - H - 1), min(d, b + 2 + H + 1)) if b > a else (a, a)) for (a, b) in zs]
            e_lo, e_hi = needs[r]
            if peer is not None:
                sq_ext = torch.empty((e_hi - e_lo, h, w), dtype=torch.int32, device=slab_u8.device)
                if e_hi > e_lo:
                    peer.pull_planes(comm, sq_ext, e_lo, e_hi)
            else:
                sq_ext = gather_planes(comm, sq, zs, needs)
            mark("halo")
            if z1 > z0:
                ridge_ext = stages.ridge(sq_ext)
                # the outermost fetched planes have an incomplete neighbourhood unless they are volume faces; they lie
                # beyond reach of [s_lo, s_hi) and must not paint
                if e_lo > 0:
                    ridge_ext[0] = 0
                if e_hi < d:
                    ridge_ext[-1] = 0
                mark("ridge")
                if sink is not None and hasattr(stages, "paint_range_into"):
                    # chunks of planes: paint one, finish the one before it (the cleanup reads two painted planes either
                    # side), hand that to the sink
                    quarter = -(-(s_hi - s_lo) // 4)
                    chunk = max(32, -(-quarter // 32) * 32)                 # about four chunks, whole front blocks of 32 planes
                    rsq_sub = stages.scratch("rsq_sub", (s_hi - s_lo, h, w), torch.int32)
                    out = torch.empty((z1 - z0, h, w), dtype=torch.float32, device=slab_u8.device)
                    part = PartialStats(0, 0.0, 0.0, float("inf"), float("-inf"))
                    done = z0
                    for c0 in range(s_lo, s_hi, chunk):
                        c1 = min(s_hi, c0 + chunk)
                        stages.paint_range_into(ridge_ext, c0 - e_lo, c1 - e_lo, gmax, rsq_sub[c0 - s_lo:c1 - s_lo])
                        mark("paint")
                        upto = z1 if c1 == s_hi else min(z1, c1 - 2)
                        if upto > done:
                            part = add_partials(part, stages.finish_range_into(rsq_sub, slab_u8, z0 - s_lo, done - s_lo, upto - s_lo,
                                                                               inverse, mask, pixel_width, out[done - z0:upto - z0]))
                            mark("finish")
                            sink(done - z0, upto - z0, out[done - z0:upto - z0])
                            done = upto
                    streamed_out = (out, part)
                else:
                    rsq_sub = stages.paint_range(ridge_ext, s_lo - e_lo, s_hi - e_lo, gmax)
                    mark("paint")
                del ridge_ext
            else:
                rsq_sub = stages.full((0, h, w), 0)
            del sq_ext
        if streamed_out is not None:
            out, part = streamed_out
        else:
            out, part = stages.finish(rsq_sub, slab_u8, z0 - s_lo, z1 - s_lo, inverse, mask, pixel_width)
            mark("finish")
            if sink is not None and z1 > z0:
                sink(0, z1 - z0, out)
    except Exception as e:                                             # noqa: BLE001 -- re-raised on every rank below
        failure = e
        out, part = None, PartialStats(0, 0.0, 0.0, float("inf"), float("-inf"))
    # (with a PeerExchange this collective is also what lets the next run overwrite sqt: every rank's pulls precede it)
    combined, nfailed = comm.combine_stats(part, slab_u8.device, failed=failure is not None)
    if failure is not None:
        raise failure
    if nfailed:
        raise RuntimeError(f"thickness_slab: {nfailed} other rank(s) failed in the ridge / paint / finish stages of this run")
    stats = combined.finalize()
    mark("stats")
    return out, stats


def map_hashes(map_f32, z0, dims, parts=8):
    """64-bit position-weighted checksums of the map's bits over the z ranges split_ranges(d, parts) that lie inside
    this slab (NaN canonicalised): sums of (bits + 1) * (2 * global index + 1) mod 2^64, so the checksums of a range
    are the same whichever rank computes them and ranks can be compared with a single-GPU run range by range.
    Returns {range index: checksum as int} for the ranges fully inside [z0, z0 + dz)."""
    w, h, d = dims
    dz = map_f32.shape[0]
    out = {}
    if dz == 0:
        return out
    bits = torch.where(torch.isnan(map_f32), torch.full_like(map_f32, float("nan")), map_f32).contiguous().view(torch.int32)
    for k, (a, b) in enumerate(split_ranges(d, parts)):
        if b <= a or a < z0 or b > z0 + dz:
            continue
        acc = 0
        for c0 in range(a, b, 16):                                     # 16 planes at a time: bounded temporaries
            c1 = min(b, c0 + 16)
            v = bits[c0 - z0:c1 - z0].reshape(-1).to(torch.int64) + 1
            idx = torch.arange(c0 * h * w, c1 * h * w, dtype=torch.int64, device=map_f32.device)
            acc = (acc + int((v * (2 * idx + 1)).sum().item())) & 0xFFFFFFFFFFFFFFFF
        out[k] = acc
    return out


# ------------------------------------------------------------------------------------------------------
# bench.py --gpus N (N > 1)

def bench(args, rank, world, local_rank, dist_mod):
    """the N > 1 arm of bench.py: returns rank 0's result fields (bench.py adds roofline, clocks and prints)"""
    from . import _native
    from .phantom import phantom_shapes

    n = args.size
    dims = (n, n, n)
    dev = torch.device("cuda", local_rank)
    ctx = _native.Context(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    comm = Comm()
    stages = CudaStages(ctx, dev)
    zs = split_ranges(n, world)
    z0, z1 = zs[rank]
    dz = z1 - z0
    shapes = phantom_shapes(n, n, n, 1)
    slab = torch.empty((dz, n, n), dtype=torch.uint8, device=dev)
    ctx.dev_phantom(shapes, n, n, n, z0, dz, slab.data_ptr())

    peer = None
    if n % world == 0 and __import__("os").environ.get("BJT_NO_PEER", "0") != "1":
        peer = PeerExchange(ctx, comm, dims, dev)

    def step(timer=None):
        th = thickness_slab(stages, comm, slab, dims, False, True, 1.0, timer, peer)
        sp = thickness_slab(stages, comm, slab, dims, True, True, 1.0, timer, peer)
        return th, sp

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize(); comm.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    timer = PhaseTimer(dev)
    launches0 = ctx.launch_count()
    ev0.record(stream)
    for _ in range(args.steps):
        th, sp = step(timer)
    ev1.record(stream)
    torch.cuda.synchronize(); comm.barrier()
    launches = ctx.launch_count() - launches0
    ms = torch.tensor([ev0.elapsed_time(ev1) / args.steps], dtype=torch.float64, device=dev)
    dist_mod.all_reduce(ms, op=dist_mod.ReduceOp.MAX)
    dev_ms = float(ms.item())
    per = {k: v / args.steps for k, v in timer.collect().items()}

    # the maps of every rank, range by range, as checksums a single-GPU run reproduces (bench.py prints the same
    # ranges at N = 1): N-GPU == 1-GPU bit for bit is checked where the numbers are compared, not just the statistics
    mine = {"tb_th": map_hashes(th[0], z0, dims), "tb_sp": map_hashes(sp[0], z0, dims)}
    gathered = [None] * world
    dist_mod.all_gather_object(gathered, mine)
    hashes = {}
    for name in ("tb_th", "tb_sp"):
        merged = {}
        for part in gathered:
            merged.update(part[name])
        hashes[name] = [format(merged[k], "016x") if k in merged else None for k in range(8)]

    # end to end: page-locked host slab in, both map slabs back out, every step
    h_in = torch.empty((dz, n, n), dtype=torch.uint8, pin_memory=True)
    h_th = torch.empty((dz, n, n), dtype=torch.float32, pin_memory=True)
    h_sp = torch.empty((dz, n, n), dtype=torch.float32, pin_memory=True)
    h_in.copy_(slab)
    slab2 = torch.empty_like(slab)
    e2e = []
    copy_stream = torch.cuda.Stream(device=dev)                     # the first map goes home while the second run computes
    for i in range(args.warmup + args.steps):
        torch.cuda.synchronize(); comm.barrier()
        t0 = time.perf_counter()
        slab2.copy_(h_in, non_blocking=True)

        def sink_to(host):
            def sink(a, b, view):                                   # finished planes go home on the copy stream
                copy_stream.wait_stream(stream)
                with torch.cuda.stream(copy_stream):
                    host[a:b].copy_(view, non_blocking=True)
            return sink
        th = thickness_slab(stages, comm, slab2, dims, False, True, 1.0, None, peer, sink_to(h_th))
        sp = thickness_slab(stages, comm, slab2, dims, True, True, 1.0, None, peer, sink_to(h_sp))
        torch.cuda.synchronize(); comm.barrier()
        if i >= args.warmup:
            e2e.append(time.perf_counter() - t0)
    em = torch.tensor([1e3 * sum(e2e) / len(e2e)], dtype=torch.float64, device=dev)
    dist_mod.all_reduce(em, op=dist_mod.ReduceOp.MAX)
    e2e_ms = float(em.item())
    # per-phase device time: the slowest rank of every phase (what the step waits for)
    names = sorted(per)
    pt = torch.tensor([per[k] for k in names], dtype=torch.float64, device=dev)
    dist_mod.all_reduce(pt, op=dist_mod.ReduceOp.MAX)
    per_max = {k: float(v) for k, v in zip(names, pt.tolist())}
    if peer is not None:
        torch.cuda.synchronize(); comm.barrier()
        peer.close()

    # The same volume through ONE process: bjt_create_multi over all the GPUs of the job and the blocking C-ABI call a
    # JNI shim binds (bjt_thickness_both_u8, page-locked host buffers) -- how the reference's single-call seam
    # (ThicknessWrapper.java:127-134, 190-196) reaches N GPUs.  Rank 0 drives it while the other ranks wait on the
    # host (a gloo barrier: an NCCL barrier would keep their GPUs spinning).
    multi = None
    if not getattr(args, "no_multi", False) and n <= 1024:
        host_group = dist_mod.new_group(backend="gloo")
        th, sp = (None, th[1]), (None, sp[1])                               # keep the statistics, drop the maps
        del slab2, h_in, h_th, h_sp
        torch.cuda.empty_cache()
        if rank == 0:
            import ctypes
            import numpy as np
            L = _native.lib()
            N = n ** 3
            vol = ctx.phantom(shapes, n, n, n)
            ctx.release_workspace()
            p_in, p_th, p_sp = L.bjt_host_alloc(N), L.bjt_host_alloc(4 * N), L.bjt_host_alloc(4 * N)
            ctypes.memmove(p_in, vol.ctypes.data, N)
            del vol
            with _native.Context(devices=list(range(world))) as mctx:
                times = []
                for i in range(2 + args.steps):
                    t0 = time.perf_counter()
                    (s0, _), (s1, _) = mctx.thickness_both_raw(p_in, n, n, n, True, 128, 1.0, p_th, p_sp)
                    if i >= 2:
                        times.append(time.perf_counter() - t0)
                ms_multi = 1e3 * sum(times) / len(times)
                mm = {}
                for name, ptr in (("tb_th", p_th), ("tb_sp", p_sp)):
                    a = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_float)), shape=(n, n, n))
                    hh = {}
                    for k, (za, zb) in enumerate(split_ranges(n, 8)):           # the same ranges and checksum as map_hashes
                        t = torch.from_numpy(a[za:zb]).to(dev)
                        hh.update(map_hashes(t, za, dims, 8))
                        del t
                    mm[name] = [format(hh[k], "016x") for k in range(8)]
                multi = dict(e2e_ms=ms_multi, hashes=mm, count=[int(s0.count), int(s1.count)], mean=[s0.mean, s1.mean])
            for ptr in (p_in, p_th, p_sp):
                L.bjt_host_free(ptr)
        dist_mod.barrier(group=host_group)
    return dict(multi=multi, dev_ms=dev_ms, e2e_ms=e2e_ms, launches=int(launches), layout_changes="peer pulls (CUDA IPC)" if peer is not None else "nccl send/recv", phases_rank0_ms={k: round(v, 3) for k, v in per.items()},
                phases_max_ms={k: round(v, 3) for k, v in per_max.items()}, hashes=hashes,
                results={"tb_th": th[1], "tb_sp": sp[1]})
