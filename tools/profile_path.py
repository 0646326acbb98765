"""ncu target: 3 E_cost evaluations of L96 D=8192, M=1025 on the path given by MFVGP_PATH."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from helpers import make_free_energy, synthetic_l96_device
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1025
g = synthetic_l96_device(torch, 8192, M, seed=8192)
fe = make_free_energy(g)
for _ in range(3): F, grad = fe.E_cost(g["x0"])
torch.cuda.synchronize(); print(F.item())
