"""torchrun target: sharded E_cost of L96 D=8192, M=4097 -- ms per evaluation, kernel ms, checksum."""
import ctypes, os, sys, time
from pathlib import Path
import torch, torch.distributed as dist
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from meanfieldvargp_b200 import _lib
from meanfieldvargp_b200.parallel import ShardedFreeEnergy
from meanfieldvargp_b200.synthetic import synthetic_l96_device
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4097
g = synthetic_l96_device(torch, 8192, M, seed=8192)
fe = ShardedFreeEnergy(g, rank, world)
x = g["x0"]
for _ in range(3):
    F, grad = fe.E_cost(x)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 20
e0.record()
for _ in range(steps):
    F, grad = fe.E_cost(x)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
lib = _lib.load()
tot, cnt = ctypes.c_double(), ctypes.c_int64()
lib.mfvgp_timing(fe.local._h, 1, ctypes.byref(tot), ctypes.byref(cnt))
for _ in range(5):
    fe.E_cost(x)
lib.mfvgp_timing(fe.local._h, 0, ctypes.byref(tot), ctypes.byref(cnt))
st = fe.local.check_last()
if rank == 0:
    print(f"MGBENCH world={world} M={M} xchg={fe.local.describe().get('xchg')} ms_per_eval={ms.item():.4f} "
          f"evals_per_s={1e3 / ms.item():.1f} kernel_ms={tot.value / max(cnt.value, 1):.4f} F={F.item():.10f} status={st}", flush=True)
dist.barrier(); dist.destroy_process_group()
