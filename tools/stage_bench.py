"""per-stage device times at n^3, device-resident (kernel iteration helper): python tools/stage_bench.py [n] [reps] [stages]
stages: comma list of edt,ridge,paint,finish (default all).  Prints one JSON line per run (Tb.Th, Tb.Sp)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bonej2_b200 import _native
from bonej2_b200.phantom import phantom_shapes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
want = set((sys.argv[3] if len(sys.argv) > 3 else "edt,ridge,paint,finish").split(","))
ctx = _native.Context(0)
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream)
N = n ** 3
d_in = torch.empty(N, dtype=torch.uint8, device="cuda")
ctx.dev_phantom(phantom_shapes(n, n, n, 1), n, n, n, 0, n, d_in.data_ptr())
gx = torch.empty(N, dtype=torch.int16, device="cuda")
a = torch.empty(N, dtype=torch.int32, device="cuda")
b = torch.empty(N, dtype=torch.int32, device="cuda")
out = torch.empty(N, dtype=torch.float32, device="cuda")


def timed(fn):
    best = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); torch.cuda.synchronize()
        best.append(e0.elapsed_time(e1))
    best.sort()
    return round(best[len(best) // 2], 3)


for inv in (False, True):
    r = {"n": n, "inverse": inv}
    fx = lambda: ctx.dev_edt_x(d_in.data_ptr(), n, n, n, inv, 128, gx.data_ptr())
    fy = lambda: ctx.dev_edt_y(gx.data_ptr(), n, n, n, n, a.data_ptr())
    fz = lambda: ctx.dev_edt_z(a.data_ptr(), n, n, n, n, b.data_ptr())
    fx(); fy(); fz()
    if "edt" in want:
        r["edt_x"], r["edt_y"], r["edt_z"] = timed(fx), timed(fy), timed(fz)
    fr = lambda: ctx.dev_ridge(b.data_ptr(), n, n, n, a.data_ptr())
    fr()
    if "ridge" in want:
        r["ridge"] = timed(fr)
    fp = lambda: ctx.dev_paint(a.data_ptr(), n, n, n, b.data_ptr())
    fp()
    if "paint" in want:
        r["paint"] = timed(fp)
    ff = lambda: ctx.dev_finish_range(b.data_ptr(), d_in.data_ptr(), n, n, n, 0, n, inv, True, 1.0, out.data_ptr())
    st = ff()
    if "finish" in want:
        r["finish"] = timed(ff)
    r["mean"] = st.mean
    print(json.dumps(r), flush=True)
