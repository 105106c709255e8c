"""paint paths side by side on one phantom (GPU): python tools/paint_compare.py N inverse [paths...]
prints per path the painter's time and whether the painted r^2 equals path 0's"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bonej2_b200 import _native
from bonej2_b200.phantom import phantom_shapes
n = int(sys.argv[1]); inv = int(sys.argv[2]); paths = [int(p) for p in sys.argv[3:]] or [0, 4]
ctx = _native.Context(0)
dev = torch.device("cuda:0")
vol = torch.empty((n, n, n), dtype=torch.uint8, device=dev)
ctx.dev_phantom(phantom_shapes(n, n, n, 1), n, n, n, 0, n, vol.data_ptr())
gx = torch.empty((n, n, n), dtype=torch.int16, device=dev)
a = torch.empty((n, n, n), dtype=torch.int32, device=dev)
sq = torch.empty((n, n, n), dtype=torch.int32, device=dev)
ctx.dev_edt_x(vol.data_ptr(), n, n, n, inv, 128, gx.data_ptr())
ctx.dev_edt_y(gx.data_ptr(), n, n, n, n, a.data_ptr())
ctx.dev_edt_z(a.data_ptr(), n, n, n, n, sq.data_ptr())
ridge = a
ctx.dev_ridge(sq.data_ptr(), n, n, n, ridge.data_ptr())
del gx
ref = None
for p in paths:
    out = torch.empty((n, n, n), dtype=torch.int32, device=dev)
    for rep in range(2):
        torch.cuda.synchronize(); t = time.time()
        ctx.dev_paint(ridge.data_ptr(), n, n, n, out.data_ptr(), path=p)
        torch.cuda.synchronize(); t = time.time() - t
    same = None if ref is None else bool(torch.equal(out, ref))
    if ref is None: ref = out
    print(f"path {p}: {t*1e3:.1f} ms, equals first path: {same}, checksum {int(out.to(torch.int64).sum())}", flush=True)
