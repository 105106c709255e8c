"""z-slab sharded Thickness: one process per GPU, torch.distributed for the plumbing.

The volume [d][h][w] is cut into contiguous z-slabs, one per rank.  Per pipeline run:

  EDT x, y          slab-local kernels
  EDT z             all-to-all transpose z-slabs -> y-slabs, z pass on whole z lines, transpose back
  ridge             on the slab's own planes (one fetched plane of squared EDT either side)
  paint             the painter's x and y stages are plane-local; its z sweeps cross slab faces.  What crosses
                    is, per (x, y) column, the staircase (tag, reach beyond the face) of the balls centred
                    within r_max planes of the face (r_max from an all-reduce of the largest squared
                    distance): every rank exports that to the slabs within reach, imports theirs and sweeps
                    its own planes (the "max-radius halo exchange" of BASELINE.json's north_star, with the
                    halo reduced to what actually reaches across).  Fallback, and the only path of backends
                    without the stepwise painter: fetch the squared-EDT planes within r_max + 2 and paint
                    slab + halo locally.
  cleanup .. stats  on slab +- 2 painted planes (fetched), output and statistics restricted to the slab
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

    def combine_stats(self, st: PartialStats, device) -> PartialStats:
        if self.size == 1:
            return st
        t = torch.tensor([float(st.count), st.sum, st.sum2, st.min, st.max], dtype=torch.float64, device=device)
        all_t = [torch.empty_like(t) for _ in range(self.size)]
        dist.all_gather(all_t, t, group=self.group)
        rows = [x.cpu().tolist() for x in all_t]                      # rank order: the same sum on every rank
        return PartialStats(int(sum(r[0] for r in rows)), sum(r[1] for r in rows), sum(r[2] for r in rows),
                            min(r[3] for r in rows), max(r[4] for r in rows))

    def barrier(self):
        if self.size > 1:
            dist.barrier(group=self.group)


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

    def _i32(self, shape):
        return torch.empty(shape, dtype=torch.int32, device=self.device)

    def edt_x(self, slab_u8, inverse, threshold=128):
        dz, h, w = slab_u8.shape
        gx = torch.empty((dz, h, w), dtype=torch.int16, device=self.device)       # u16 bits
        self.ctx.dev_edt_x(slab_u8.data_ptr(), w, h, dz, inverse, threshold, gx.data_ptr())
        return gx

    def edt_y(self, gx, n_sentinel):
        dz, h, w = gx.shape
        gy = self._i32((dz, h, w))
        self.ctx.dev_edt_y(gx.data_ptr(), w, h, dz, n_sentinel, gy.data_ptr())
        return gy

    def edt_z(self, gy_t, n_sentinel):
        d, hy, w = gy_t.shape
        sq = self._i32((d, hy, w))
        if hy > 0:
            self.ctx.dev_edt_z(gy_t.data_ptr(), w, hy, d, n_sentinel, sq.data_ptr())
        return sq

    def max_value(self, t):
        return int(t.max().item()) if t.numel() else 0

    def ridge(self, sq_ext):
        de, h, w = sq_ext.shape
        out = self._i32((de, h, w))
        self.ctx.dev_ridge(sq_ext.data_ptr(), w, h, de, out.data_ptr())
        return out

    def paint(self, ridge_ext):
        de, h, w = ridge_ext.shape
        out = self._i32((de, h, w))
        self.ctx.dev_paint(ridge_ext.data_ptr(), w, h, de, out.data_ptr())
        return out

    # stepwise painter (bjt_dev_paint_open .. _close); boundary buffers are int32 tensors [ncol * (2 kcap + 1)]:
    # items first (8-byte aligned), counts behind
    def paint_open(self, ridge_own, max_rsq):
        dz, h, w = ridge_own.shape
        self._ps = (ridge_own, self._i32((dz, h, w)))
        return self.ctx.dev_paint_open(ridge_own.data_ptr(), w, h, dz, max_rsq)

    def paint_xslab_cols(self, xs):
        return self.ctx.dev_paint_xslab_width(xs) * self._ps[0].shape[1]

    def paint_block_cols(self, shape):
        """column-block width the library would choose for a slab of this shape (not clipped to the width)"""
        dz, h, w = shape
        return self.ctx.paint_block_cols(w, h, dz)

    def paint_set_block_cols(self, cols):
        self.ctx.paint_set_block_cols(cols)

    def paint_lists(self, xs):
        self.ctx.dev_paint_lists(xs)

    def boundary_buffer(self, ncol, kcap):
        return torch.empty(ncol * (2 * kcap + 1), dtype=torch.int32, device=self.device)

    @staticmethod
    def _split(buf, ncol, kcap):
        return buf.data_ptr() + 8 * ncol * kcap, buf.data_ptr()              # (count pointer, items pointer)

    def paint_export(self, xs, side, gap, nplanes, buf, ncol, kcap):
        cnt, items = self._split(buf, ncol, kcap)
        self.ctx.dev_paint_export(xs, side, gap, nplanes, cnt, items, kcap)

    def paint_sweep(self, xs, lo, hi, ncol, kcap):
        self.ctx.dev_paint_sweep(xs, [self._split(b, ncol, kcap) for b in lo], [self._split(b, ncol, kcap) for b in hi],
                                 kcap, self._ps[1].data_ptr())

    def paint_close(self):
        flags = self.ctx.dev_paint_close()
        out = self._ps[1]
        self._ps = None
        return out, flags

    def plane_weights(self, ridge):
        """cost estimate per plane of a ridge slab (see plane_costs)"""
        nz, h, w = ridge.shape
        q = ridge.sum(dim=(1, 2), dtype=torch.float64) / float(h * w)
        return (float(h * w) * (COST_BASE + q * q)).tolist()

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


# ------------------------------------------------------------------------------------------------------

def no_result(n: int) -> int:
    return 3 * (n + 1) * (n + 1)


def balanced_ranges(weights, parts: int):
    """contiguous ranges of range(len(weights)) with near-equal weight sums (ranges may be empty)"""
    n = len(weights)
    total = float(sum(weights))
    if n == 0 or total <= 0.0:
        return split_ranges(n, parts)
    cuts, acc, k = [0], 0.0, 1
    for i, wgt in enumerate(weights):
        acc += float(wgt)
        while k < parts and acc >= k * total / parts:
            cuts.append(i + 1)
            k += 1
    while len(cuts) < parts:
        cuts.append(n)
    cuts.append(n)
    return [(cuts[i], max(cuts[i], cuts[i + 1])) for i in range(parts)]


# Painter cost of a plane, from its ridge points.  Measured on 128-plane slabs of the 1024^3 phantom (tools/slab_cost.py):
# Tb.Th 18-19 ms everywhere; Tb.Sp 45-54 ms inside the volume but 72 and 83 ms for the two face slabs, whose balls are
# larger.  With q = (sum of the ridge r^2 of the plane) / (voxels of the plane), cost ~ voxels x (113 + q^2) reproduces the
# three levels (q = 1.5, 13.4, 19.5 -> 115 : 293 : 493 against 18.6 : 47 : 83 ms).  Only the ratios matter, and a poor
# estimate costs balance, never correctness.
COST_BASE = 113.0


def plane_costs(comm: Comm, stages, ridge_own, zs):
    """estimated painter cost of every plane of the volume (the same list on every rank)"""
    mine = stages.plane_weights(ridge_own) if ridge_own.shape[0] else []
    if comm.size == 1:
        return list(mine)
    parts = [None] * comm.size
    dist.all_gather_object(parts, [float(x) for x in mine], group=comm.group)
    return [x for part in parts for x in part]


NO_BLOCK_PREFERENCE = 1 << 20   # column-block width a rank without planes proposes (a multiple of 32 above any width)
def even_block_cols(w: int, limit: int, quantum: int = 128) -> int:
    """column-block width <= limit that cuts a width of w into blocks of (nearly) equal size: 1024 columns under a
    limit of 896 become 2 x 512, not 896 + 128 (a narrow last block leaves the GPU mostly idle)"""
    if limit >= w:
        return limit
    nblocks = -(-w // limit)
    even = -(-(-(-w // nblocks)) // quantum) * quantum
    return min(limit, even)


BOUNDARY_KCAP = 32          # staircase steps per column a slab face can carry (more: fall back to the halo repaint)
MAX_SOURCES = 4             # slabs per side a sweep can import from (bjt_dev_paint_sweep)


def paint_exchange(stages, comm: Comm, ridge_own, zs, gmax, kcap=BOUNDARY_KCAP, timings=None):
    """Painted r^2 of this rank's own planes: plane-local x / y stages, boundary staircases exchanged with the
    slabs within reach, z sweeps.  Returns None (on every rank alike) when the stepwise painter cannot take
    the case -- too many slabs within reach or a capacity exceeded -- and the caller repaints slab + halo."""
    G, r = comm.size, comm.rank
    z0, z1 = zs[r]
    rmax = math.isqrt(gmax)                                               # no ball reaches further than this
    nplanes = rmax + 1

    def within(p, q):
        """gap from slab p's face plane to slab q's first plane, or 0 when out of reach / empty"""
        (a0, a1), (b0, b1) = zs[p], zs[q]
        if a1 == a0 or b1 == b0:
            return 0
        gap = b0 - (a1 - 1) if q > p else a0 - (b1 - 1)
        return gap if gap <= rmax else 0

    too_many = any(sum(1 for p in range(q) if within(p, q)) > MAX_SOURCES or
                   sum(1 for p in range(q + 1, G) if within(p, q)) > MAX_SOURCES for q in range(G))
    if too_many:
        return None
    # Stages 2-3 run per block of columns and staircases are exchanged per block, so every rank must cut the
    # width the same way -- but the library sizes the blocks by the slab's own volume, and slabs differ in
    # thickness (uneven division, the cost-balanced split): agree on the narrowest choice.
    agreed = hasattr(stages, "paint_block_cols")
    if agreed:
        mine = stages.paint_block_cols(tuple(ridge_own.shape)) if z1 > z0 else NO_BLOCK_PREFERENCE
        narrowest = -comm.allreduce_max_int(-int(mine), ridge_own.device)
        stages.paint_set_block_cols(even_block_cols(int(ridge_own.shape[2]), narrowest))
    flags = 0
    out = None
    t_ = time.perf_counter

    def mark(name, t0):
        if timings is not None:
            if str(getattr(stages, "device", "cpu")).startswith("cuda"):
                torch.cuda.synchronize()
            timings[name] = timings.get(name, 0.0) + (t_() - t0)
        return t_()

    try:
        if z1 > z0:
            t0 = t_()
            nxs = stages.paint_open(ridge_own, gmax)
            t0 = mark("paint.x_stage", t0)
            for xs in range(nxs):
                ncol = stages.paint_xslab_cols(xs)
                stages.paint_lists(xs)
                t0 = mark("paint.y_stage", t0)
                sends, recvs, lo, hi = {}, {}, [], []
                for p in range(G):
                    if p == r:
                        continue
                    gap = within(r, p)
                    if gap:                                                   # what I carry across to slab p
                        buf = stages.boundary_buffer(ncol, kcap)
                        stages.paint_export(xs, +1 if p > r else -1, gap, nplanes, buf, ncol, kcap)
                        sends[p] = buf
                    if within(p, r):                                          # what slab p carries across to me
                        recvs[p] = stages.boundary_buffer(ncol, kcap)
                        (lo if p < r else hi).append(recvs[p])
                t0 = mark("paint.export", t0)
                comm.exchange(sends, recvs)
                t0 = mark("paint.exchange", t0)
                stages.paint_sweep(xs, lo, hi, ncol, kcap)
                t0 = mark("paint.z_stage", t0)
                del sends, recvs, lo, hi
            out, flags = stages.paint_close()
    finally:
        if agreed:
            stages.paint_set_block_cols(0)
    flags = comm.allreduce_max_int(int(flags != 0), ridge_own.device)
    if flags:
        return None
    return out if out is not None else ridge_own.new_empty((0,) + tuple(ridge_own.shape[1:]))


def thickness_slab(stages, comm: Comm, slab_u8, dims, inverse, mask=True, pixel_width=1.0, timings=None,
                   paint_mode="auto", boundary_kcap=BOUNDARY_KCAP, balance=True):
    """One pipeline run (LocalThicknessWrapper.processImage + StackStatistics) on this rank's z-slab.
    slab_u8: [dz][h][w] planes zs[rank] of the volume dims = (w, h, d).  Returns (map slab f32, stats dict)."""
    w, h, d = dims
    G, r = comm.size, comm.rank
    zs, ys = split_ranges(d, G), split_ranges(h, G)
    z0, z1 = zs[r]
    n = max(w, h, d)
    n0 = no_result(n)
    t_ = time.perf_counter

    def mark(name, t0):
        if timings is not None:
            if hasattr(stages, "device") and str(stages.device).startswith("cuda"):
                torch.cuda.synchronize()
            timings[name] = timings.get(name, 0.0) + (t_() - t0)

    t0 = t_()
    gx = stages.edt_x(slab_u8, inverse)
    gy = stages.edt_y(gx, n)
    del gx
    mark("edt_xy", t0); t0 = t_()
    gy_t = transpose_z_to_y(comm, gy, zs, ys)
    del gy
    mark("transpose", t0); t0 = t_()
    sq_t = stages.edt_z(gy_t, n)
    del gy_t
    mark("edt_z", t0); t0 = t_()
    sq = transpose_y_to_z(comm, sq_t, zs, ys, h)
    del sq_t
    mark("transpose", t0); t0 = t_()
    gmax = comm.allreduce_max_int(stages.max_value(sq), slab_u8.device)

    # planes of the painted map the cleanup of this slab reads: slab +- 2, clipped to the volume
    s_lo, s_hi = max(0, z0 - 2), min(d, z1 + 2)
    if gmax == 0:
        rsq_sub = stages.full((s_hi - s_lo, h, w), 0)
    elif gmax >= n0:
        rsq_sub = stages.full((s_hi - s_lo, h, w), n0)                 # no background anywhere (sentinel)
    else:
        rsq_sub = None
        if paint_mode != "halo" and hasattr(stages, "paint_open"):
            # ridge of the slab's own planes: the squared EDT one plane further out on either side
            needs = [(max(0, a - 1), min(d, b + 1)) if b > a else (a, a) for (a, b) in zs]
            sq_ext = gather_planes(comm, sq, zs, needs)
            mark("halo", t0); t0 = t_()
            ridge_own = stages.ridge(sq_ext)[z0 - needs[r][0]:z1 - needs[r][0]] if z1 > z0 else sq_ext
            del sq_ext
            mark("ridge", t0); t0 = t_()
            # the painter works on its own split of the planes: equal estimated cost instead of equal thickness (balls are
            # larger near the volume faces and bone is not spread evenly); ridge planes move over, painted planes move back
            zp = zs
            if balance and G > 1 and hasattr(stages, "plane_weights"):
                zp = balanced_ranges(plane_costs(comm, stages, ridge_own, zs), G)
                if zp != zs:
                    ridge_own = gather_planes(comm, ridge_own, zs, zp)
                mark("rebalance", t0); t0 = t_()
            rsq_own = paint_exchange(stages, comm, ridge_own, zp, gmax, kcap=boundary_kcap, timings=timings)
            del ridge_own
            mark("paint", t0); t0 = t_()
            if rsq_own is not None:
                needs2 = [(max(0, a - 2), min(d, b + 2)) if b > a else (a, a) for (a, b) in zs]
                rsq_sub = gather_planes(comm, rsq_own, zp, needs2)
                del rsq_own
                mark("halo", t0); t0 = t_()
            elif paint_mode == "exchange":
                raise RuntimeError("stepwise painter refused the case (capacity or reach) and paint_mode='exchange' forbids the fallback")
    if gmax != 0 and gmax < n0 and rsq_sub is None:
        H = math.isqrt(gmax - 1) + 1 + 2                               # ceil(sqrt(r^2 max)) + 2
        # ridge of planes [z0-H, z1+H) needs the squared EDT one plane further out
        needs = []
        for (a, b) in zs:
            needs.append((max(0, a - H - 1), min(d, b + H + 1)))
        e_lo, e_hi = needs[r]
        sq_ext = gather_planes(comm, sq, zs, needs)
        mark("halo", t0); t0 = t_()
        ridge_ext = stages.ridge(sq_ext)
        del sq_ext
        # the outermost fetched planes have an incomplete neighbourhood unless they are volume faces;
        # they lie beyond reach (H + 1 > r_max + 2) and are dropped
        if e_lo > 0:
            ridge_ext[0] = 0
        if e_hi < d:
            ridge_ext[-1] = 0
        mark("ridge", t0); t0 = t_()
        rsq_ext = stages.paint(ridge_ext)
        del ridge_ext
        rsq_sub = rsq_ext[s_lo - e_lo:s_hi - e_lo]
        mark("paint", t0); t0 = t_()
    out, part = stages.finish(rsq_sub, slab_u8, z0 - s_lo, z1 - s_lo, inverse, mask, pixel_width)
    stats = comm.combine_stats(part, slab_u8.device).finalize()
    mark("finish", t0)
    return out, stats


# ------------------------------------------------------------------------------------------------------
# bench.py --gpus N (N > 1)

def bench(args, rank, world, local_rank, dist_mod):
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

    def step(timings=None):
        th = thickness_slab(stages, comm, slab, dims, False, True, 1.0, timings)
        sp = thickness_slab(stages, comm, slab, dims, True, True, 1.0, timings)
        return th, sp

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize(); comm.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    timings = {}
    ev0.record(stream)
    for _ in range(args.steps):
        th, sp = step(timings)
    ev1.record(stream)
    torch.cuda.synchronize(); comm.barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1) / args.steps], dtype=torch.float64, device=dev)
    dist_mod.all_reduce(ms, op=dist_mod.ReduceOp.MAX)
    dev_ms = float(ms.item())

    # end to end: page-locked host slab in, both map slabs back out, every step
    nb = dz * n * n
    h_in = torch.empty((dz, n, n), dtype=torch.uint8, pin_memory=True)
    h_th = torch.empty((dz, n, n), dtype=torch.float32, pin_memory=True)
    h_sp = torch.empty((dz, n, n), dtype=torch.float32, pin_memory=True)
    h_in.copy_(slab)
    slab2 = torch.empty_like(slab)
    launches0 = ctx.launch_count()
    e2e = []
    copy_stream = torch.cuda.Stream(device=dev)                     # the first map goes home while the second run computes
    for i in range(args.warmup + args.steps):
        torch.cuda.synchronize(); comm.barrier()
        t0 = time.perf_counter()
        slab2.copy_(h_in, non_blocking=True)
        th = thickness_slab(stages, comm, slab2, dims, False, True, 1.0)
        copy_stream.wait_stream(stream)
        with torch.cuda.stream(copy_stream):
            h_th.copy_(th[0], non_blocking=True)
        th[0].record_stream(copy_stream)
        sp = thickness_slab(stages, comm, slab2, dims, True, True, 1.0)
        h_sp.copy_(sp[0], non_blocking=True)
        torch.cuda.synchronize(); comm.barrier()
        if i >= args.warmup:
            e2e.append(time.perf_counter() - t0)
    launches = (ctx.launch_count() - launches0) // (args.warmup + args.steps)
    em = torch.tensor([1e3 * sum(e2e) / len(e2e)], dtype=torch.float64, device=dev)
    dist_mod.all_reduce(em, op=dist_mod.ReduceOp.MAX)
    e2e_ms = float(em.item())

    N = n ** 3
    per = {k: 1e3 * v / args.steps for k, v in timings.items()}
    dom = max(per, key=per.get) if per else "paint"
    alg_edt = 21.0 * N * 2 / world
    edt_ms = per.get("edt_xy", 0.0) + per.get("edt_z", 0.0)
    return {
        "metric": "thickness_map_gvoxels_per_s", "value": N / (dev_ms * 1e-3) / 1e9, "unit": "Gvoxel/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u32/f32", "data": "synthetic",
        "config": {"workload": f"BoneJ Thickness Tb.Th+Tb.Sp (mapChoice Both), mask on, {n}^3 synthetic trabecular phantom, "
                               f"z-slabs over {world} GPUs (BASELINE.json configs[2])", "volume": [n, n, n],
                   "parallelism": f"z-slab x{world}", "l2_policy": "inputs exceed the 126 MB L2"},
        "e2e": {"value": N / (e2e_ms * 1e-3) / 1e9, "unit": "Gvoxel/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": N, "d2h_bytes_per_step": 8 * N},
        "gpu_launches": int(launches) * args.steps,
        "roofline": {"kernel": "edt (x+y+z passes, rank 0, per step)", "bound": "hbm",
                     "achieved": alg_edt / (edt_ms * 1e-3) / 1e9 if edt_ms > 0 else 0.0, "peak": None, "unit": "GB/s",
                     "frac": None, "traffic": None},
        "stages_rank0_ms": {k: round(v, 3) for k, v in per.items()}, "dominant_stage": dom,
        "results": {"tb_th": th[1], "tb_sp": sp[1]},
        "cpu_baseline": None,
    }
