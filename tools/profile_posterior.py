"""ncu target: posterior reconstruction at D=8192, M=4097, num=20."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import meanfieldvargp_b200 as mf
D, M, num = 8192, 4097, 20
obs = 0.25 * np.arange(1, M + 1)
mp = torch.randn((D, 3 * M + 4), dtype=torch.float64, device="cuda")
sp = 4.0 + 2.0 * torch.rand((D, 2 * M + 3), dtype=torch.float64, device="cuda")
for _ in range(3):
    tt, mt, st = mf.reconstruct(mp, sp, 0.0, obs, obs[-1] + 0.25, num)
torch.cuda.synchronize(); print(float(mt[5, 77]))
