#!/usr/bin/env python
"""bench.py -- BoneJ Thickness hot path on B200: thickness-map Gvoxels/s (BASELINE.json metric).

A step is one Thickness run with mapChoice = "Both" (Tb.Th then Tb.Sp, mask on) over one synthetic
trabecular phantom -- what ThicknessWrapper.run() does for one dataset (ThicknessWrapper.java:123-156).

  value      device-only: the u8 volume is resident in HBM when the timed region starts, both float
             maps and their statistics are left in HBM; CUDA events on the stream the kernels use.
  e2e        the same work through the C-ABI call a JNI shim binds (bjt_thickness_both_u8) with
             page-locked HOST buffers: upload and both map downloads are inside the timed region.
  roofline   the dominant kernel (the painter's disc kernel): SURVEY.md 8(d)'s algorithmic bytes of one launch
             over its launch time from events inside the library, against the measured HBM peak
             (`structure_frac`: the same with the bytes the kernel's own data structure moves); every
             stage, the EDT passes included (the "EDT HBM GB/s" of the metric), is under `stages`.
  map_hash   position-weighted 64-bit checksums of both maps over eight z ranges: the same at every N when
             the N-GPU maps equal the single-GPU maps bit for bit.
  cpu_baseline  the CPU oracle (a restatement of sc.fiji LocalThickness_, no JVM in the image) on
             a bounded sample of the same workload, on the box's host cores.

N > 1 (torchrun): the volume is split into z-slabs, one per rank (bonej2_b200/sharded.py).
--impl reference times the CPU oracle only (rank 0), same metric and config.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "thickness_map_gvoxels_per_s"
UNIT = "Gvoxel/s"
SEED = 1


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_sample(n_sample, threads):
    """oracle on an n_sample^3 phantom of the same recipe, Both, mask on -> (Gvoxel/s, seconds)"""
    import numpy as np  # noqa: F401
    from bonej2_b200.phantom import phantom_shapes
    from oracle import lt_oracle
    lt_oracle.build()
    vol = lt_oracle.phantom_raster(phantom_shapes(n_sample, n_sample, n_sample, SEED), n_sample, n_sample, n_sample)
    t0 = time.perf_counter()
    for inverse in (False, True):
        lt_oracle.process_image(vol, inverse=inverse, mask=True, nthreads=threads)
    dt = time.perf_counter() - t0
    return (n_sample ** 3) / dt / 1e9, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_sample = args.cpu_size
    times = []
    for i in range(args.warmup + args.steps):
        g, dt = cpu_sample(n_sample, threads)
        if i >= args.warmup:
            times.append(dt)
    dt = sum(times) / len(times)
    value = (n_sample ** 3) / dt / 1e9
    sample = f"{n_sample}^3 phantom (same recipe, seed {SEED}), Both, mask on; whole pipeline incl. brute-force paint"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u32/f32", "data": "synthetic",
        "config": {"workload": f"BoneJ Thickness Tb.Th+Tb.Sp (mapChoice Both), mask on, synthetic trabecular phantom: "
                               f"{n_sample}^3 timed here (bounded CPU sample of the {args.size}^3 workload of the GPU arm)",
                   "volume_timed": [n_sample] * 3, "volume_gpu_arm": [args.size] * 3, "timed_sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "CPU restatement of sc.fiji LocalThickness_ (no JVM in image)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def run_hyperstack(args, rank, world, local_rank, dist):
    """BASELINE.json configs[4]: a multi-channel / time hyperstack, one Thickness "Both" run per {X, Y, Z} subspace in
    HyperstackUtils.split3DSubspaces order (HyperstackUtils.java:99-136).  C3 x T2 of --hs-size^3 phantoms (seeds 1..5)
    with one all-background subspace (NaN Tb.Th row, constant Tb.Sp).  N = 1: the subspaces one after the other on one
    GPU; N > 1: subspace k on rank k mod N (independent: no data-path collective, only the rows are gathered).
    value = subspaces resident in HBM, maps left in HBM; e2e = bonej2_b200.hyperstack's public call on host arrays."""
    import numpy as np
    import torch
    from bonej2_b200 import _native, hyperstack
    from bonej2_b200.phantom import phantom_shapes
    from bonej2_b200.wrapper import CHANNEL, TIME, X, Y, Z
    n, C_, T_ = args.hs_size, 3, 2
    nsub = C_ * T_
    ctx = _native.Context(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    vols = []
    for k in range(nsub):
        if k == nsub - 1:
            vols.append(np.zeros((n, n, n), np.uint8))                       # the empty subspace
        else:
            vols.append(ctx.phantom(phantom_shapes(n, n, n, SEED + k), n, n, n))
    data = np.stack(vols).reshape(T_, C_, n, n, n)                           # numpy order T, C, Z, Y, X
    axes = [X, Y, Z, CHANNEL, TIME]
    mine = [k for k in range(nsub) if k % world == rank]
    d_in = [torch.from_numpy(vols[k]).cuda() for k in mine]
    d_out = torch.empty(n ** 3, dtype=torch.float32, device="cuda")

    def dev_step():
        nl = 0
        for t in d_in:
            for inverse in (False, True):
                _, tm = ctx.dev_thickness(t.data_ptr(), n, n, n, inverse=inverse, mask=True, d_out=d_out.data_ptr())
                nl += tm.n_launches
        return nl

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        dev_step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    launches = 0
    for _ in range(args.steps):
        launches += dev_step()
    ev1.record(stream)
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1) / args.steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dev_ms = float(ms.item())
    # end to end: the module's public call, host arrays in, rows (and maps) out
    e2e = []
    rows = None
    for i in range(args.warmup + args.steps):
        barrier()
        t0 = time.perf_counter()
        if world > 1:
            rows = hyperstack.thickness_hyperstack_ranks(data, axes, name="hyper", map_choice="Both", device=local_rank)
        else:
            rows, _maps = hyperstack.thickness_hyperstack(data, axes, name="hyper", map_choice="Both", devices=(local_rank,),
                                                          want_maps=True)
        barrier()
        if i >= args.warmup:
            e2e.append(time.perf_counter() - t0)
    em = torch.tensor([1e3 * sum(e2e) / len(e2e)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(em, op=dist.ReduceOp.MAX)
    e2e_ms = float(em.item())
    clocks = sampler.stop()
    ctx.close()
    if rank != 0:
        return
    peaks, peak_kind = measured_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    Nvox = nsub * n ** 3
    edt_note = "per-subspace stage times are those of the volume workload at this size"
    maps_back = world == 1
    line = {
        "metric": METRIC, "value": Nvox / (dev_ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32/f32", "data": "synthetic",
        "config": {"workload": f"BoneJ Thickness per subspace of a C{C_} x T{T_} hyperstack of {n}^3 synthetic trabecular phantoms "
                               f"(one empty), mapChoice Both, mask on; subspace k on GPU k mod {world} (BASELINE.json configs[4])",
                   "subspaces": nsub, "subspace_volume": [n, n, n], "parallelism": f"subspace x{world}",
                   "l2_policy": "every subspace (16.8 MB u8, 67 MB intermediates at 256^3) is rewritten between its uses by the "
                                "other subspaces' runs; inputs of one step exceed the 126 MB L2 together"},
        "e2e": {"value": Nvox / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": Nvox,
                "d2h_bytes_per_step": 8 * Nvox if maps_back else 0,
                "call": "bonej2_b200.hyperstack.thickness_hyperstack" + ("_ranks (rows only)" if world > 1 else " (rows and maps)")},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "disc_kernel (painter, paint path 4)", "bound": "hbm", "achieved": None, "peak": hbm_peak, "unit": "GB/s",
                     "frac": None, "traffic": None, "peak_kind": peak_kind, "note": edt_note},
        "rows": [[label, {k: (None if v != v else v) for k, v in vals.items()}] for label, vals in rows],
        "cpu_baseline": None, "clocks": clocks,
    }
    emit(line)


_REAL_STDOUT = None


def emit(obj):
    """the one JSON line, on the process's original stdout"""
    data = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # libraries print to stdout as they please (NCCL announces its version there); the contract is ONE JSON
    # line, so everything else that reaches fd 1 is sent to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=1024, help="phantom edge in voxels (the same volume at every N: strong scaling)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-size", type=int, default=256, help="edge of the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="volume", choices=["volume", "hyperstack"],
                    help="volume: one --size^3 phantom (BASELINE.json configs[2] / [3]); hyperstack: configs[4], a C3 x T2 "
                         "stack of --hs-size^3 phantoms, one of them empty, one Thickness run per subspace")
    ap.add_argument("--hs-size", type=int, default=256, help="edge of one subspace of the hyperstack workload")
    ap.add_argument("--no-multi", action="store_true", help="N > 1: skip the one-process bjt_create_multi leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200" and not os.environ.get("BJT_BENCH_SHORT_WARMUP"):
        args.warmup = 3                     # timing rule: at least 3 untimed steps (the profiling scripts set the variable)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    from bonej2_b200 import _native
    from bonej2_b200.phantom import phantom_shapes

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the Thickness path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if args.workload == "hyperstack":
        run_hyperstack(args, rank, world, local_rank, dist)
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_kind = measured_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))

    if world > 1:
        from bonej2_b200 import sharded
        sampler = ClockSampler(local_rank)
        sampler.start()
        r = sharded.bench(args, rank, world, local_rank, dist)
        clocks = sampler.stop()
        if rank == 0:
            n = args.size
            N = n ** 3
            ph = r["phases_max_ms"]
            paint_ms = ph.get("paint", 0.0)
            # the same dominant kernel as at N = 1 (the painter's disc kernel is inside `paint`; the phase also holds the
            # front kernel): SURVEY.md 8(d) bytes of the painter over the slowest rank's paint phase
            alg_paint = 4.0 * N * 2                                   # + 16 B per ridge point, not gathered here: a lower bound
            gbs = alg_paint / world / (paint_ms * 1e-3) / 1e9 if paint_ms > 0 else 0.0
            edt_ms = ph.get("edt_xy", 0.0) + ph.get("edt_z", 0.0)
            edt_gbs = 21.0 * N * 2 / world / (edt_ms * 1e-3) / 1e9 if edt_ms > 0 else 0.0
            line = {
                "metric": METRIC, "value": N / (r["dev_ms"] * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": r["dev_ms"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "u32/f32", "data": "synthetic",
                "config": {"workload": f"BoneJ Thickness Tb.Th+Tb.Sp (mapChoice Both), mask on, {n}^3 synthetic trabecular "
                                       f"phantom, z-slabs over {world} GPUs (BASELINE.json configs[{3 if n >= 2048 else 2}])",
                           "volume": [n, n, n], "parallelism": f"z-slab x{world}", "seed": SEED,
                           "l2_policy": "inputs exceed the 126 MB L2"},
                "e2e": {"value": N / (r["e2e_ms"] * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": r["e2e_ms"],
                        "h2d_bytes_per_step": N, "d2h_bytes_per_step": 8 * N},
                "gpu_launches": r["launches"],
                "roofline": {"kernel": "disc_kernel (painter, paint path 4: sparse-table disc tiles); here with the front kernel: the "
                                       "slowest rank's whole paint phase per step",
                             "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                             "traffic": None, "peak_kind": peak_kind,
                             "note": "per GPU; instruction-issue bound, not HBM bound (profiles/README.md)"},
                "stages": {"edt_total": {"ms_per_step": round(edt_ms, 3), "GBps_per_gpu": round(edt_gbs, 1),
                                         "frac_of_hbm_peak": round(edt_gbs / hbm_peak, 4)}},
                "phases_max_over_ranks_ms": ph, "phases_rank0_ms": r["phases_rank0_ms"],
                "map_hash": r["hashes"], "results": r["results"], "cpu_baseline": None, "clocks": clocks,
                "layout_changes": r.get("layout_changes"),
            }
            if r.get("multi"):
                m = r["multi"]
                line["e2e_one_process"] = {
                    "value": N / (m["e2e_ms"] * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": m["e2e_ms"], "h2d_bytes_per_step": N,
                    "d2h_bytes_per_step": 8 * N, "call": f"bjt_create_multi({world} devices) + bjt_thickness_both_u8, page-locked host buffers",
                    "maps_match_ranks": m["hashes"] == r["hashes"], "map_hash": m["hashes"]}
            if not args.no_cpu:
                threads = os.cpu_count() or 1
                g, dt = cpu_sample(args.cpu_size, threads)
                line["cpu_baseline"] = {"value": g, "unit": UNIT, "cores": threads, "kind": "port",
                                        "sample": f"{args.cpu_size}^3 phantom (same recipe, seed {SEED}), Both, mask on, {dt:.1f} s",
                                        "note": "CPU restatement of sc.fiji LocalThickness_ (no JVM in image)"}
            emit(line)
        dist.destroy_process_group()
        return

    n = args.size
    N = n ** 3
    ctx = _native.Context(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    info = ctx.device_info()

    # synthetic phantom, rasterised on the device straight into the resident input buffer
    shapes = phantom_shapes(n, n, n, SEED)
    d_in = torch.empty(N, dtype=torch.uint8, device="cuda")
    ctx.dev_phantom(shapes, n, n, n, 0, n, d_in.data_ptr())
    bvtv = float((d_in != 0).float().mean().item())
    d_th = torch.empty(N, dtype=torch.float32, device="cuda")
    d_sp = torch.empty(N, dtype=torch.float32, device="cuda")

    def dev_step():
        s0, t0 = ctx.dev_thickness(d_in.data_ptr(), n, n, n, inverse=False, mask=True, d_out=d_th.data_ptr())
        s1, t1 = ctx.dev_thickness(d_in.data_ptr(), n, n, n, inverse=True, mask=True, d_out=d_sp.data_ptr())
        return (s0, t0), (s1, t1)

    for _ in range(args.warmup):
        dev_step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms = {}
    launches = 0
    n_ridge = [0, 0]
    y_sweeps = 0
    torch.cuda.synchronize()
    ev0.record(stream)
    for _ in range(args.steps):
        res = dev_step()
        for k, (st, tm) in enumerate(res):
            d = tm.as_dict()
            launches += d["n_launches"]
            n_ridge[k] = d["n_ridge"]
            for key in ("ms_edt_x", "ms_edt_y", "ms_edt_z", "ms_ridge", "ms_paint", "ms_finish",
                        "ms_paint_x", "ms_paint_y", "ms_paint_z"):
                stage_ms[key] = stage_ms.get(key, 0.0) + d[key]
            y_sweeps += d["n_paint_y_sweeps"]
    ev1.record(stream)
    torch.cuda.synchronize()
    dev_ms = ev0.elapsed_time(ev1) / args.steps
    stats_dev = [res[0][0].as_dict(), res[1][0].as_dict()]
    paint_paths = [res[0][1].paint_path, res[1][1].paint_path]
    from bonej2_b200.sharded import map_hashes
    hashes = {}
    for name, t in (("tb_th", d_th), ("tb_sp", d_sp)):
        hh = map_hashes(t.view(n, n, n), 0, (n, n, n))
        hashes[name] = [format(hh[k], "016x") if k in hh else None for k in range(8)]

    # ---- end to end through the C-ABI with page-locked host buffers --------------------------------
    L = _native.lib()
    h_in = L.bjt_host_alloc(N)
    h_th = L.bjt_host_alloc(4 * N)
    h_sp = L.bjt_host_alloc(4 * N)
    if not (h_in and h_th and h_sp):
        raise SystemExit("could not allocate page-locked host buffers")
    import ctypes
    torch.cuda.synchronize()
    host_vol = d_in.cpu().numpy()                                  # host copy of the same phantom
    ctypes.memmove(h_in, host_vol.ctypes.data, N)
    del host_vol
    e2e_times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        (s0, tm0), (s1, tm1) = ctx.thickness_both_raw(h_in, n, n, n, True, 128, 1.0, h_th, h_sp)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            e2e_times.append(dt)
            launches_e2e = tm0.n_launches + tm1.n_launches
    clocks = sampler.stop()
    e2e_ms = 1e3 * sum(e2e_times) / len(e2e_times)
    # The path a JVM drives (bonej2_b200/csrc/jni_shim.cpp): one pageable array per slice, copied slice by slice into the
    # page-locked staging buffers (Get<Type>ArrayRegion is a memcpy out of the Java heap), the same C-ABI call, and the
    # maps copied back slice by slice (Set<Type>ArrayRegion).  The arrays here are numpy's, the copies the same memcpy.
    h_in_v = np.ctypeslib.as_array(ctypes.cast(h_in, ctypes.POINTER(ctypes.c_uint8)), shape=(n, n * n))
    h_th_v = np.ctypeslib.as_array(ctypes.cast(h_th, ctypes.POINTER(ctypes.c_float)), shape=(n, n * n))
    h_sp_v = np.ctypeslib.as_array(ctypes.cast(h_sp, ctypes.POINTER(ctypes.c_float)), shape=(n, n * n))
    sl_in = [h_in_v[z].copy() for z in range(n)]
    sl_th = [np.empty(n * n, np.float32) for _ in range(n)]
    sl_sp = [np.empty(n * n, np.float32) for _ in range(n)]
    slice_times = []
    for i in range(1 + min(args.steps, 3)):
        t0 = time.perf_counter()
        for z in range(n):
            np.copyto(h_in_v[z], sl_in[z])
        ctx.thickness_both_raw(h_in, n, n, n, True, 128, 1.0, h_th, h_sp)
        for z in range(n):
            np.copyto(sl_th[z], h_th_v[z])
            np.copyto(sl_sp[z], h_sp_v[z])
        if i >= 1:
            slice_times.append(time.perf_counter() - t0)
    e2e_slices_ms = 1e3 * sum(slice_times) / len(slice_times)
    del sl_in, sl_th, sl_sp, h_in_v, h_sp_v
    # the host maps must agree with the device-resident ones
    del h_th_v
    th_host = np.ctypeslib.as_array(ctypes.cast(h_th, ctypes.POINTER(ctypes.c_float)), shape=(N,))
    chk = slice(0, N, max(1, N // (1 << 22)))
    same = bool(np.array_equal(np.nan_to_num(th_host[chk], nan=-1.0), np.nan_to_num(d_th[chk].cpu().numpy(), nan=-1.0)))
    L.bjt_host_free(h_in); L.bjt_host_free(h_th); L.bjt_host_free(h_sp)

    # ---- roofline ------------------------------------------------------------------------------------
    # algorithmic bytes per voxel and pipeline run: EDT 21 (x 5, y 8, z 8), ridge 8,
    # 2*sqrt + cleanup 16 and mask/NaN/calibrate/stats 9 (fused here: "finish" 25), paint 4 + 16 per ridge point.
    per_step = {k: v / args.steps for k, v in stage_ms.items()}      # ms per step, both runs together
    alg = {"ms_edt_x": 5.0 * N * 2, "ms_edt_y": 8.0 * N * 2, "ms_edt_z": 8.0 * N * 2, "ms_ridge": 8.0 * N * 2,
           "ms_finish": 25.0 * N * 2, "ms_paint": 4.0 * N * 2 + 16.0 * (n_ridge[0] + n_ridge[1])}
    stages = {}
    for k, a in alg.items():
        ms = per_step[k]
        gbs = a / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        stages[k[3:]] = {"ms_per_step": round(ms, 3), "alg_GB": round(a / 1e9, 3), "GBps": round(gbs, 1),
                         "frac_of_hbm_peak": round(gbs / hbm_peak, 4)}
    for k in ("ms_paint_x", "ms_paint_y", "ms_paint_z"):
        stages[k[3:]] = {"ms_per_step": round(per_step[k], 3)}
    edt_ms = per_step["ms_edt_x"] + per_step["ms_edt_y"] + per_step["ms_edt_z"]
    edt_gbs = 21.0 * N * 2 / (edt_ms * 1e-3) / 1e9
    stages["edt_total"] = {"ms_per_step": round(edt_ms, 3), "alg_GB": round(42.0 * N / 1e9, 3), "GBps": round(edt_gbs, 1),
                           "frac_of_hbm_peak": round(edt_gbs / hbm_peak, 4)}
    # The dominant kernel.  Paint path 4: disc_kernel (one launch per chunk of planes; ms_paint_y = their sum from events
    # inside the library, on the stream the kernels run on).  `achieved` uses SURVEY.md 8(d)'s algorithmic bytes of the
    # painter (4 B per voxel written + 16 B per ridge point) -- that figure belongs to the painter as a whole, so it is
    # divided by the whole painter's time and attributed per launch of the disc kernel by its share.  `structure_*`:
    # the bytes this kernel's own data structure moves per launch (items read once per tile they touch are not counted
    # twice: 8 B per front item + 8 B descriptor per 32 voxels read, 4 B per voxel written).
    sweeps_per_step = y_sweeps / args.steps
    if sweeps_per_step > 0 and per_step["ms_paint_y"] > 0:
        disc = paint_paths[1] == 4
        launch_ms = per_step["ms_paint_y"] / sweeps_per_step
        alg_launch = alg["ms_paint"] / sweeps_per_step
        gbs = alg["ms_paint"] / (per_step["ms_paint"] * 1e-3) / 1e9
        if disc:
            struct_launch = (4.0 + 0.25) * N * 2 / sweeps_per_step + 8.0 * 1.6 * N * 2 / sweeps_per_step
            name = "disc_kernel (painter, paint path 4: sparse-table disc tiles)"
        else:
            struct_launch = 24.0 * N * 4.0 / sweeps_per_step
            name = "sep_dyn_kernel<PACKED,2,+-1> (painter y-stage sweep)"
        traffic = traffic_build = None
        try:
            rec = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "disc_traffic.json")))
            if rec.get("volume") == [n, n, n] and disc:
                traffic = rec["dram_bytes_per_launch"]
                traffic_build = rec.get("build")
        except (OSError, ValueError, KeyError):
            pass
        roofline = {"kernel": name, "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                    "traffic": traffic, "peak_kind": peak_kind, "launch_ms": round(launch_ms, 3),
                    "launches_per_step": sweeps_per_step, "algorithmic_bytes_per_launch": alg_launch,
                    "structure_bytes_per_launch": struct_launch,
                    "structure_frac": struct_launch / (launch_ms * 1e-3) / 1e9 / hbm_peak,
                    "share_of_step": round(per_step["ms_paint_y"] / dev_ms, 4),
                    "note": "`achieved` = SURVEY.md 8(d) painter bytes over the whole painter's time; the kernel is bound by "
                            "instruction issue and shared-memory atomics, not HBM (profiles/README.md); `stages` holds the "
                            "streaming kernels"}
        if traffic_build:
            roofline["traffic_capture"] = traffic_build
    else:   # no list-based painter ran (constant fast path / scatter fallback): report the slowest stage
        dom = max(alg, key=lambda k: per_step[k])
        dom_gbs = alg[dom] / (per_step[dom] * 1e-3) / 1e9
        roofline = {"kernel": dom[3:], "bound": "hbm", "achieved": dom_gbs, "peak": hbm_peak, "unit": "GB/s",
                    "frac": dom_gbs / hbm_peak, "traffic": None, "peak_kind": peak_kind}

    # ---- CPU baseline (bounded sample) -----------------------------------------------------------------
    cpu = None
    if not args.no_cpu:
        threads = os.cpu_count() or 1
        g, dt = cpu_sample(args.cpu_size, threads)
        cpu = {"value": g, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{args.cpu_size}^3 phantom (same recipe, seed {SEED}), Both, mask on, {dt:.1f} s",
               "note": "CPU restatement of sc.fiji LocalThickness_ (no JVM in image)"}

    line = {
        "metric": METRIC, "value": N / (dev_ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32/f32", "data": "synthetic",
        "config": {"workload": f"BoneJ Thickness Tb.Th+Tb.Sp (mapChoice Both), mask on, {n}^3 synthetic trabecular phantom "
                               f"(BASELINE.json configs[2] at 1 GPU)", "volume": [n, n, n], "bv_tv": round(bvtv, 4),
                   "seed": SEED, "l2_policy": "inputs (1 GiB u8, 4 GiB intermediates) exceed the 126 MB L2",
                   "paint_path": paint_paths, "device": info["name"], "sm_count": info["sm_count"]},
        "e2e": {"value": N / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": N,
                "d2h_bytes_per_step": 8 * N, "host_maps_match_device": same},
        "e2e_slices": {"value": N / (e2e_slices_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e2e_slices_ms,
                       "note": "as a JVM drives it: one pageable array per slice, copied through the page-locked staging "
                               "buffers slice by slice in and out (what jni_shim.cpp does with Get/Set<Type>ArrayRegion), "
                               "single host thread; the copies of 9 bytes per voxel through host memory dominate"},
        "gpu_launches": int(launches),
        "roofline": roofline, "stages": stages, "map_hash": hashes, "cpu_baseline": cpu, "clocks": clocks,
        "results": {"tb_th": {k: stats_dev[0][k] for k in ("count", "mean", "sd", "max")},
                    "tb_sp": {k: stats_dev[1][k] for k in ("count", "mean", "sd", "max")}},
    }
    emit(line)


if __name__ == "__main__":
    main()
