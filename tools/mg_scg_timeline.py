"""torchrun target: CUPTI kernel timeline of one sharded SCG iteration (L96 D=8192, M=4097), rank 0."""
import contextlib, io, os, sys
from pathlib import Path
import torch, torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from meanfieldvargp_b200.parallel import ShardedFreeEnergy
from meanfieldvargp_b200.synthetic import synthetic_l96_device
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4097
g = synthetic_l96_device(torch, 8192, M, seed=8192)
fe = ShardedFreeEnergy(g, rank, world)
x = g["x0"]
with contextlib.redirect_stdout(io.StringIO()):
    fe.find_minimum(x, max_iter=10)
torch.cuda.synchronize(); dist.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    with contextlib.redirect_stdout(io.StringIO()):
        fe.find_minimum(x, max_iter=10)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.end)
rows = [(e.name.split("(")[0].split("::")[-1][:34], e.time_range.start, e.time_range.end) for e in ev]
walks = [i for i, r in enumerate(rows) if "esde_l96_walk" in r[0]]
if rank == 0 and len(walks) >= 10:
    i0, i1 = walks[6], walks[8]
    t0 = rows[i0][1]
    prev = None
    for name, st, en in rows[i0:i1 + 1]:
        gap = "" if prev is None else f"idle-before {max(0.0, st - prev):7.1f}"
        print(f"  {name:36s} start {st - t0:9.1f} end {en - t0:9.1f} dur {en - st:8.1f}  {gap}")
        prev = en
    print(f"  one iteration (two evaluations): {rows[i1][1] - rows[i0][1]:.1f} us")
dist.barrier(); dist.destroy_process_group()
