"""torchrun target: CUPTI kernel timeline (torch.profiler) of sharded E_cost evaluations of
L96 D=8192 on every rank: start, duration and gap of the kernels of one evaluation."""
import os, sys
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
for _ in range(5):
    fe.E_cost(x)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(8):
        fe.E_cost(x)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.end)          # PDL successors become resident early: order by END
rows = [(e.name.split("(")[0].split("::")[-1][:34], e.time_range.start, e.time_range.end - e.time_range.start) for e in ev]
walks = [i for i, r in enumerate(rows) if "esde_l96_walk" in r[0]]
lines = []
if len(walks) >= 6:
    i0, i1 = walks[4], walks[5]
    t0 = rows[i0][1]
    prev_end = None
    for name, st, du in rows[i0:i1 + 1]:
        gap = "" if prev_end is None else f"end-to-end {st + du - prev_end:6.2f}"
        lines.append(f"  {name:36s} start {st - t0:9.2f} us  end {st + du - t0:9.2f} us  dur {du:8.2f} us  {gap}")
        prev_end = st + du
    lines.append(f"  period {rows[i1][1] - rows[i0][1]:.2f} us")
for r in range(world):
    if r == rank:
        print(f"rank {rank} intervals {fe.local.intervals} xchg={fe.local.describe().get('xchg')}")
        print("\n".join(lines), flush=True)
    dist.barrier()
dist.destroy_process_group()
