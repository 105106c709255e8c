"""per-phase timings of the z-slab sharded path on slabs of chosen thickness (torchrun, N ranks):
   python -m torch.distributed.run --nproc-per-node N tools/slab_bench.py W H D [steps]
Ranks take consecutive slabs of the first D planes of the W x H x max(W,H,D)... phantom (same recipe as bench.py)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from bonej2_b200 import _native, sharded
from bonej2_b200.phantom import phantom_shapes

w, h, d = (int(x) for x in sys.argv[1:4])
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
ctx = _native.Context(lr)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
comm, stages = sharded.Comm(), sharded.CudaStages(ctx, dev)
zs = sharded.split_ranges(d, world)
z0, z1 = zs[rank]
full = max(w, h, d, 1024)
slab = torch.empty((z1 - z0, h, w), dtype=torch.uint8, device=dev)
ctx.dev_phantom(phantom_shapes(w, h, full, 1), w, h, full, z0, z1 - z0, slab.data_ptr())
for it in range(1 + steps):
    tm = {}
    torch.cuda.synchronize(); comm.barrier()
    t0 = time.perf_counter()
    for inv in (False, True):
        sharded.thickness_slab(stages, comm, slab, (w, h, d), inv, True, 1.0, tm if it else None)
    torch.cuda.synchronize(); comm.barrier()
    dt = time.perf_counter() - t0
if True:
    rows = [None] * world
    if world > 1:
        dist.all_gather_object(rows, {k: round(1e3 * v, 1) for k, v in tm.items()})
    else:
        rows = [{k: round(1e3 * v, 1) for k, v in tm.items()}]
    if rank == 0:
        print(json.dumps({"dims": [w, h, d], "ranks": world, "ms_step_wall": round(1e3 * dt, 1), "gvox_s": round(w * h * d / dt / 1e9, 3)}))
        for r, row in enumerate(rows):
            print(r, row)
if world > 1:
    dist.destroy_process_group()
