"""painter cost of z-slabs against cheap per-slab features of the ridge (for the work-balanced split of sharded runs):
   python tools/slab_cost.py [N] [slabs]"""
import sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import torch
from bonej2_b200 import _native, sharded
from bonej2_b200.phantom import phantom_shapes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
G = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda", 0)
ctx = _native.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
st = sharded.CudaStages(ctx, dev)
vol = torch.empty((n, n, n), dtype=torch.uint8, device=dev)
ctx.dev_phantom(phantom_shapes(n, n, n, 1), n, n, n, 0, n, vol.data_ptr())
for inverse in (False, True):
    gy = st.edt_y(st.edt_x(vol, inverse), n)
    sq = st.edt_z(gy, n); del gy
    gmax = int(sq.max().item())
    ridge = st.ridge(sq); del sq
    zs = sharded.split_ranges(n, G)
    rows = []
    for (a, b) in zs:
        r = ridge[a:b]
        rf = r[r > 0].to(torch.float32)
        feats = (float(rf.numel()), float(rf.sqrt().sum()), float(rf.sum()))
        for rep in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            nxs = st.paint_open(r, gmax)
            tx = None
            torch.cuda.synchronize(); t1 = time.perf_counter()
            for xs in range(nxs):
                st.paint_lists(xs)
            torch.cuda.synchronize(); t2 = time.perf_counter()
            for xs in range(nxs):
                st.paint_sweep(xs, [], [], st.paint_xslab_cols(xs), 32)
            out, flags = st.paint_close()
            torch.cuda.synchronize(); t3 = time.perf_counter()
        rows.append((a, b, feats, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)))
        del out
    print("inverse", inverse, "gmax", gmax)
    for a, b, f, tx, ty, tz in rows:
        print(f"  planes {a:4d}-{b:4d}  n_ridge {f[0]:.3e}  sum r {f[1]:.3e}  sum r2 {f[2]:.3e}   x {tx:6.1f}  y {ty:6.1f}  z {tz:6.1f}  total {tx+ty+tz:6.1f} ms")
    del ridge
