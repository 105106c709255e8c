"""per-run stage times of the device-resident pipeline (bjt_dev_thickness_u8), median of reps: python tools/pipeline_runs.py [n] [reps]"""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bonej2_b200 import _native
from bonej2_b200.phantom import phantom_shapes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ctx = _native.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
d_in = torch.empty(n ** 3, dtype=torch.uint8, device="cuda")
ctx.dev_phantom(phantom_shapes(n, n, n, 1), n, n, n, 0, n, d_in.data_ptr())
d_out = torch.empty(n ** 3, dtype=torch.float32, device="cuda")
keys = ("ms_edt_x", "ms_edt_y", "ms_edt_z", "ms_ridge", "ms_paint", "ms_paint_x", "ms_paint_y", "ms_finish", "ms_total")
for inv in (False, True):
    rows = []
    for _ in range(reps + 1):
        st, tm = ctx.dev_thickness(d_in.data_ptr(), n, n, n, inverse=inv, mask=True, d_out=d_out.data_ptr())
        rows.append(tm.as_dict())
    print(json.dumps({"inverse": inv, **{k[3:]: round(statistics.median(r[k] for r in rows[1:]), 3) for k in keys}}), flush=True)
