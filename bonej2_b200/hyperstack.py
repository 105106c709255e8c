"""Hyperstack front-end for batched Thickness (BASELINE.json configs[4]).

`ThicknessWrapper` itself refuses images with channel or time axes (ThicknessWrapper.java:236-243); the
other BoneJ wrappers handle them by splitting into {X, Y, Z} subspaces with
HyperstackUtils.split3DSubspaces (Modern/wrapperPlugins/.../wrapperUtils/HyperstackUtils.java:99-136,
291-320) and labelling result rows `name + " " + subspace` (e.g. SurfaceAreaWrapper.java:136-144).  This
module does the same split -- first split axis fastest, labels as Subspace.toString (:398-433) with
ResultUtils.toConventionalIndex (1-based for Z, Channel, Time) -- and runs the CUDA Thickness path once
per subspace, optionally spreading subspaces over several GPUs (they are independent).  There is no CPU
path: each subspace goes through the C-ABI like any other volume.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass

import numpy as np

from . import _native
from .wrapper import CHANNEL, TIME, UNKNOWN, X, Y, Z

AXIS_NAMES = {X: "X", Y: "Y", Z: "Z", CHANNEL: "Channel", TIME: "Time"}


def _conventional_index(axis_type, index: int) -> int:
    """ResultUtils.toConventionalIndex (ResultUtils.java:201-209)"""
    return index + 1 if axis_type in (Z, CHANNEL, TIME) else index


@dataclass
class Subspace:
    data: np.ndarray            # [z][y][x] view (contiguous copy)
    positions: tuple            # position along each split axis, in split-axis order
    axis_types: tuple
    label: str                  # Subspace.toString(): "Channel: 1, Time: 2"


def split_3d_subspaces(data: np.ndarray, axes, axis_names=None):
    """Yield the {X, Y, Z} subspaces of an N-d array in HyperstackUtils order.

    data : numpy array, slowest axis first (numpy order), i.e. data.shape == reversed(sizes of `axes`)
    axes : axis types, first axis fastest (imglib2 order), e.g. [X, Y, CHANNEL, Z, TIME]
    Axes of the same type get a subscript "(2)", "(3)" ... as in HyperstackUtils.mapTypeSubscripts.
    """
    axes = list(axes)
    n = len(axes)
    if data.ndim != n:
        raise ValueError("one axis type per array dimension")
    spatial = [i for i, a in enumerate(axes) if a in (X, Y, Z)]
    if not spatial:
        return
    split = [i for i in range(n) if i not in spatial]
    sizes = [data.shape[n - 1 - i] for i in range(n)]          # size of imglib2 axis i
    seen, subs = {}, []
    for i in split:
        seen[axes[i]] = seen.get(axes[i], 0) + 1
        subs.append(seen[axes[i]])
    names = axis_names or AXIS_NAMES
    # first split axis varies fastest: iterate the cartesian product with the LAST split axis outermost
    ranges = [range(sizes[i]) for i in reversed(split)]
    for combo in itertools.product(*ranges):
        pos = tuple(reversed(combo))                              # in split-axis order
        index = [slice(None)] * n
        for i, p in zip(split, pos):
            index[n - 1 - i] = p
        view = data[tuple(index)]
        if view.size == 0:
            continue
        # remaining axes are the spatial ones in imglib2 order; bring to [z][y][x]
        rem = [axes[i] for i in spatial]                          # fastest first
        order = [rem.index(t) for t in (X, Y, Z) if t in rem]     # imglib2 positions of x, y, z
        arr = np.transpose(view, [len(rem) - 1 - o for o in reversed(order)])
        while arr.ndim < 3:
            arr = arr[np.newaxis]
        parts = []
        for i, p, s in zip(split, pos, subs):
            t = axes[i]
            nm = names.get(t, "Unknown") + (f"({s})" if s > 1 else "")
            parts.append(f"{nm}: {_conventional_index(t, p)}")
        yield Subspace(np.ascontiguousarray(arr), pos, tuple(axes[i] for i in split), ", ".join(parts))


def _runs(map_choice):
    if map_choice not in ("Trabecular thickness", "Trabecular separation", "Both"):
        raise ValueError("Unexpected map choice")
    return [("Tb.Th", False)] if map_choice == "Trabecular thickness" else \
           [("Tb.Sp", True)] if map_choice == "Trabecular separation" else [("Tb.Th", False), ("Tb.Sp", True)]


def thickness_hyperstack_ranks(data, axes, name="image", map_choice="Both", mask=True, pixel_width=1.0, unit="",
                               device=None, run=None, group=None):
    """The same batch over the ranks of a torch.distributed job (one process per GPU): subspace k is handled
    by rank k mod world, every rank returns all rows in subspace order.  Subspaces are independent, so there
    is no data-path collective -- only the result rows are gathered.

    run(vol_u8, inverse) -> (mean, sd, max) may be injected (tests on CPU over gloo); by default it is the
    CUDA path on `device` (default: the rank's local device) and fails without it."""
    import torch.distributed as dist
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    runs = _runs(map_choice)
    unit_header = f"({unit or 'pixel'})"
    ctx = None
    if run is None:
        import os
        ctx = _native.Context(int(os.environ.get("LOCAL_RANK", rank)) if device is None else device)

        cache = {}

        def run(vol, inverse):
            # "Both": one call per subspace (one upload); the second run's row comes from the same call
            key = id(vol)
            if key not in cache:
                cache.clear()
                res = _run_subspace(ctx, vol, map_choice, mask, pixel_width, False)
                cache[key] = {inv: res[prefix][0] for prefix, inv in runs}
            st = cache[key][inverse]
            return (st.mean, st.sd, st.max) if st.count else (float("nan"),) * 3
    mine = []
    try:
        for k, sub in enumerate(split_3d_subspaces(np.asarray(data), axes)):
            if k % world != rank:
                continue
            if sub.data.dtype != np.uint8:
                raise ValueError("Need a binary image")
            vals = {}
            for prefix, inverse in runs:
                mean, sd, mx = run(sub.data, inverse)
                vals[f"{prefix} Mean {unit_header}"] = mean
                vals[f"{prefix} Std Dev {unit_header}"] = sd
                vals[f"{prefix} Max {unit_header}"] = mx
            mine.append((k, f"{name} {sub.label}".strip(), vals))
    finally:
        if ctx is not None:
            ctx.close()
    parts = [mine]
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, mine, group=group)
    rows = sorted((row for part in parts for row in part), key=lambda r: r[0])
    return [(label, vals) for _, label, vals in rows]


def _run_subspace(ctx, vol, map_choice, mask, pixel_width, want_maps):
    """one subspace through the C-ABI: {prefix: (stats, map)}; "Both" is one call (one upload, the first map on its way
    back while the second run computes: bjt_thickness_both_u8), as ThicknessWrapper.run does for one image"""
    if map_choice == "Both":
        (m0, s0, _), (m1, s1, _) = ctx.thickness_both(vol, mask=mask, pixel_width=pixel_width, want_maps=want_maps)
        return {"Tb.Th": (s0, m0), "Tb.Sp": (s1, m1)}
    prefix, inverse = _runs(map_choice)[0]
    m, st, _ = ctx.thickness(vol, inverse=inverse, mask=mask, pixel_width=pixel_width, want_map=want_maps)
    return {prefix: (st, m)}


def thickness_hyperstack(data, axes, name="image", map_choice="Both", mask=True, pixel_width=1.0, unit="",
                         devices=(0,), want_maps=False):
    """Thickness on every {X, Y, Z} subspace.  Returns rows [(label, {header: value})] in subspace order
    (and the maps when want_maps).  Subspaces are dealt round-robin to the given CUDA devices and the devices work
    at the same time: one host thread per device drives its context (the calls block in native code with the
    interpreter lock released), so k devices take k subspaces at once.  A failure on any device is raised after the
    others have finished their current subspace."""
    import threading
    runs = _runs(map_choice)
    unit_header = f"({unit or 'pixel'})"
    subs = list(split_3d_subspaces(np.asarray(data), axes))
    for sub in subs:
        if sub.data.dtype != np.uint8:
            raise ValueError("Need a binary image")
    ctxs = [_native.Context(d) for d in devices]
    results, errors = [None] * len(subs), []

    def worker(i):
        try:
            for k in range(i, len(subs), len(ctxs)):
                if errors:
                    return
                results[k] = _run_subspace(ctxs[i], subs[k].data, map_choice, mask, pixel_width, want_maps)
        except Exception as e:                                   # re-raised on the calling thread
            errors.append(e)

    try:
        if len(ctxs) == 1:
            worker(0)
        else:
            threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(ctxs))]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
    finally:
        for c in ctxs:
            c.close()
    if errors:
        raise errors[0]
    rows, maps = [], []
    for sub, res in zip(subs, results):
        vals, out = {}, {}
        for prefix, _ in runs:
            st, m = res[prefix]
            mean, sd, mx = (st.mean, st.sd, st.max) if st.count else (float("nan"),) * 3
            vals[f"{prefix} Mean {unit_header}"] = mean
            vals[f"{prefix} Std Dev {unit_header}"] = sd
            vals[f"{prefix} Max {unit_header}"] = mx
            out[prefix] = m
        rows.append((f"{name} {sub.label}".strip(), vals))
        maps.append(out)
    return (rows, maps) if want_maps else rows
