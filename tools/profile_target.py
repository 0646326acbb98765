"""ncu target: 3 E_cost evals of L96 D=500 (M=16), then 2 of D=8192 (M=4097)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import make_free_energy, synthetic_l96, synthetic_l96_device  # noqa: E402

g = synthetic_l96(500, 16, seed=8192)
fe = make_free_energy(g)
x = torch.tensor(g["x0"], dtype=torch.float64, device="cuda")
for _ in range(3):
    F, grad = fe.E_cost(x)
torch.cuda.synchronize()
print("D=500 F =", F.item())
big = len(sys.argv) < 2 or sys.argv[1] != "small"
if big:
    g = synthetic_l96_device(torch, 8192, 4097, seed=8192)
    fe2 = make_free_energy(g)
    for _ in range(2):
        F, grad = fe2.E_cost(g["x0"])
    torch.cuda.synchronize()
    print("D=8192 F =", F.item())
