"""one Thickness call on a synthetic phantom (profiling helper): python tools_one_run.py N inverse [reps] [paint_path]"""
import sys, time
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from bonej2_b200 import _native
from bonej2_b200.phantom import phantom_shapes
n = int(sys.argv[1]); inv = int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ctx = _native.Context(0)
if len(sys.argv) > 4: ctx.set_paint_path(int(sys.argv[4]))
import os
if os.environ.get('BJT_CAP'): ctx.debug_set_paint_limits(int(os.environ['BJT_CAP']), 0)     # experiment: send big lines to the redo kernels
vol = ctx.phantom(phantom_shapes(n, n, n, 1), n, n, n)
for _ in range(reps):
    m, st, tm = ctx.thickness(vol, inverse=bool(inv), mask=True)
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in tm.as_dict().items()}, st.mean)
