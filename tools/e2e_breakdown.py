"""where the end-to-end time of bjt_thickness_both_u8 goes (pinned host buffers): python tools/e2e_breakdown.py [N]"""
import ctypes, sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np
from bonej2_b200 import _native
from bonej2_b200.phantom import phantom_shapes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = n ** 3
ctx = _native.Context(0)
L = _native.lib()
h_in, h_th, h_sp = L.bjt_host_alloc(N), L.bjt_host_alloc(4 * N), L.bjt_host_alloc(4 * N)
vol = ctx.phantom(phantom_shapes(n, n, n, 1), n, n, n)
ctypes.memmove(h_in, vol.ctypes.data, N)
for i in range(3):
    t0 = time.perf_counter()
    (s0, t0m), (s1, t1m) = ctx.thickness_both_raw(h_in, n, n, n, True, 128, 1.0, h_th, h_sp)
    wall = 1e3 * (time.perf_counter() - t0)
    f = lambda t: {k: round(getattr(t, k), 1) for k in ("ms_h2d", "ms_edt_x", "ms_edt_y", "ms_edt_z", "ms_ridge", "ms_paint", "ms_finish", "ms_d2h", "ms_total")}
    print("wall", round(wall, 1), "Th", f(t0m), "Sp", f(t1m))
