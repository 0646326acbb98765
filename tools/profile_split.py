import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from helpers import make_free_energy, synthetic_l96_device
g = synthetic_l96_device(torch, 8192, 1025, seed=8192)
fe2 = make_free_energy(g)
for _ in range(3): F, grad = fe2.E_cost(g["x0"])
torch.cuda.synchronize(); print(F.item())
