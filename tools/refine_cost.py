"""Cost of the adaptive path at D=8192: E_cost with k steep variance triples (flagged intervals)."""
import sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from helpers import make_free_energy, synthetic_l96_device
D, M = 8192, 1025
g = synthetic_l96_device(torch, D, M, seed=8192)
fe = make_free_energy(g)
x0 = g["x0"].clone()
Nm, Ns = 3 * M + 4, 2 * M + 3
for nbad, val in ((0, 0.0), (1, 0.3), (10, 0.3), (40, 0.3), (10, 0.05)):
    x = x0.clone()
    if nbad:
        sp = x[D * Nm:].view(D, Ns)
        cols = torch.arange(nbad) * 50 + 21            # one odd (mid-interval) column per flagged interval
        sp[:, cols] = float(np.log(val))
    for _ in range(2): F, grad = fe.E_cost(x)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): F, grad = fe.E_cost(x)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 3 * 1e3
    info = fe.refine_info()
    print(f"steep columns {nbad:3d} (s={val}): {ms:8.2f} ms/eval, flagged intervals E {int((info[:,0]>0).sum())} dS {int((info[:,1]>0).sum())}, max bisections {info.max()}")
