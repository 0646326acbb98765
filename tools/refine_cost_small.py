"""Cost of the in-kernel adaptive path on the north-star size (L96 D=500, M=16)."""
import os, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from helpers import make_free_energy, synthetic_l96
D, M = 500, 16
g = synthetic_l96(D, M, seed=8192)
fe = make_free_energy(g)
Nm, Ns = 3 * M + 4, 2 * M + 3
for nbad, val in ((0, 0.0), (1, 0.3), (3, 0.3), (1, 0.05)):
    x = g["x0"].copy()
    if nbad:
        sp = x[D * Nm:].reshape(D, Ns)
        sp[:, [5 + 8 * i for i in range(nbad)]] = np.log(val)
    xd = torch.tensor(x, dtype=torch.float64, device="cuda")
    for _ in range(5): F, grad = fe.E_cost(xd)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50): F, grad = fe.E_cost(xd)
    torch.cuda.synchronize(); us = (time.perf_counter() - t0) / 50 * 1e6
    info = fe.refine_info()
    print(f"{os.environ.get('MFVGP_PATH','coop')}: steep columns {nbad} (s={val}): {us:8.1f} us/eval, flagged E {int((info[:,0]>0).sum())} dS {int((info[:,1]>0).sum())}, max bisections {info.max()}")
