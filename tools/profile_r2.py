"""ncu target of round 2: `python tools/profile_r2.py D M [reps]` -- reps x E_cost of synthetic L96
(D, M) on device pointers (default path for that size)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from meanfieldvargp_b200.synthetic import make_free_energy, synthetic_l96, synthetic_l96_device
D, M = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
if D * M > 4_000_000:
    g = synthetic_l96_device(torch, D, M, seed=8192)
    x = g["x0"]
else:
    g = synthetic_l96(D, M, seed=8192)
    x = torch.tensor(g["x0"], dtype=torch.float64, device="cuda")
fe = make_free_energy(g)
for _ in range(reps):
    F, grad = fe.E_cost(x)
torch.cuda.synchronize()
print(fe.describe(), "F =", F.item(), "status", fe.check_last())
