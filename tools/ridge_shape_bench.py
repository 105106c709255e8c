"""ridge kernel time over volume shapes (device-resident, synthetic squared-distance values): python tools/ridge_shape_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bonej2_b200 import _native
ctx = _native.Context(0)
stream = torch.cuda.current_stream(); ctx.set_stream(stream.cuda_stream)
for (w, h, d) in [(1024, 1024, 220), (2048, 2048, 128), (2048, 2048, 358), (2048, 1024, 358), (1024, 2048, 358), (2016, 2048, 358), (2048, 2040, 358)]:
    n = w * h * d
    sq = (torch.rand(n, device="cuda") * 50).to(torch.int32)
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    ts = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); ctx.dev_ridge(sq.data_ptr(), w, h, d, out.data_ptr()); e1.record(stream); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print((w, h, d), "ms", round(min(ts), 3), "Gvoxel/s", round(n / min(ts) / 1e6, 1), flush=True)
    del sq, out
