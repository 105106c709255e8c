import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from bonej2_b200 import _native
from bonej2_b200.phantom import phantom_shapes
ctx = _native.Context(0)
for n in [int(a) for a in sys.argv[1:]]:
    shapes = phantom_shapes(n, n, n, 1)
    t = time.time(); vol = ctx.phantom(shapes, n, n, n); tp = time.time() - t
    print(f"n={n} phantom {tp:.2f}s BV/TV {(vol > 0).mean():.4f}", flush=True)
    for rep in range(2):
        for inv in (False, True):
            t = time.time()
            m, st, tm = ctx.thickness(vol, inverse=inv, mask=True)
            wall = time.time() - t
            d = tm.as_dict()
            print(f"  inv={inv} wall {wall:.3f}s mean {st.mean:.4f} max {st.max:.3f} " + " ".join(f"{k}={v:.2f}" if isinstance(v, float) else f"{k}={v}" for k, v in d.items()), flush=True)
