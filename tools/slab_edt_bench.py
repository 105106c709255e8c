"""EDT pass times on the shapes one rank of a 2048^3 run over 8 GPUs works on (lines of 2048: 16-column pair tiles):
y pass on a z-slab [256][2048][2048], z pass on a y-slab [2048][256][2048].  python tools/slab_edt_bench.py [n] [ranks] [reps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bonej2_b200 import _native
from bonej2_b200.phantom import phantom_shapes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
G = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dz = n // G
ctx = _native.Context(0)
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream)
N = dz * n * n
d_in = torch.empty(N, dtype=torch.uint8, device="cuda")
ctx.dev_phantom(phantom_shapes(n, n, n, 1), n, n, n, n // 2, dz, d_in.data_ptr())      # planes n/2 .. n/2 + dz of the phantom
gx = torch.empty(N, dtype=torch.int16, device="cuda")
a = torch.empty(N, dtype=torch.int32, device="cuda")
b = torch.empty(N, dtype=torch.int32, device="cuda")


def timed(fn):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return round(ts[len(ts) // 2], 3)


for inv in (False, True):
    fx = lambda: ctx.dev_edt_x(d_in.data_ptr(), n, n, dz, inv, 128, gx.data_ptr())
    fy = lambda: ctx.dev_edt_y(gx.data_ptr(), n, n, dz, n, a.data_ptr())
    fz = lambda: ctx.dev_edt_z(a.data_ptr(), n, dz, n, n, b.data_ptr())        # the y-pass result read as a y-slab [n][dz][n]
    fx(); fy(); fz()
    r = {"n": n, "ranks": G, "inverse": inv, "edt_x": timed(fx), "edt_y": timed(fy), "edt_z": timed(fz)}
    r["GBps_yz"] = round(16.0 * N / ((r["edt_y"] + r["edt_z"]) * 1e-3) / 1e9, 1)
    print(json.dumps(r), flush=True)
