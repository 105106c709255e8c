"""One z-slab of a sharded run recomputed on ONE GPU, independently of bonej2_b200.sharded: the phantom planes of the
slab plus a margin of 3 * r_max planes either side go through the single-call pipeline (bjt_dev_thickness_u8); inside
the slab every stage sees exactly what it sees in the whole volume (nearest background, ridge neighbours, balls and
cleanup neighbours all lie within the margin), so the slab's part of the map -- and its position-weighted checksums --
must equal what the ranks of the N-GPU run produced.
usage: python tools/verify_slab.py N parts k [reference.json]     (slab k of `parts`; compares with map_hash of the JSON line)"""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bonej2_b200 import _native
from bonej2_b200.phantom import phantom_shapes
from bonej2_b200.sharded import map_hashes, split_ranges

n, parts, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ref = json.load(open(sys.argv[4])) if len(sys.argv) > 4 else None
z0, z1 = split_ranges(n, parts)[k]
ctx = _native.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
shapes = phantom_shapes(n, n, n, 1)
ok = True
for name, inverse in (("tb_th", False), ("tb_sp", True)):
    rmax = None
    margin = 64
    while True:
        a, b = max(0, z0 - margin), min(n, z1 + margin)
        vol = torch.empty((b - a, n, n), dtype=torch.uint8, device="cuda")
        ctx.dev_phantom(shapes, n, n, n, a, b - a, vol.data_ptr())
        out = torch.empty((b - a, n, n), dtype=torch.float32, device="cuda")
        st, tm = ctx.dev_thickness(vol.data_ptr(), n, n, b - a, inverse=inverse, mask=True, d_out=out.data_ptr())
        rmax = math.isqrt(tm.rsq_max) + 1
        if margin >= 3 * rmax + 4 or (a == 0 and b == n):
            break
        margin = 3 * rmax + 4                                        # the first guess was too small for this r_max: again
        del vol, out
    hh = map_hashes(out[z0 - a:z1 - a], z0, (n, n, n), parts=8)
    got = {i: format(v, "016x") for i, v in hh.items()}
    line = {"map": name, "slab": [z0, z1], "planes_computed": [a, b], "r_max": rmax, "hash": got}
    if ref is not None:
        want = {i: ref["map_hash"][name][i] for i in got}
        line["matches_reference"] = want == got
        ok &= want == got
    print(json.dumps(line), flush=True)
    del vol, out
sys.exit(0 if ok else 1)
