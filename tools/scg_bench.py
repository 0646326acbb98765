"""Device SCG iterations per second on the north-star size (L96 D=500, M=16)."""
import contextlib, io, os, sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from helpers import make_free_energy, synthetic_l96
g = synthetic_l96(500, 16, seed=8192)
fe = make_free_energy(g)
x = torch.tensor(g["x0"], dtype=torch.float64, device="cuda")
with contextlib.redirect_stdout(io.StringIO()):
    fe.find_minimum(x, max_iter=10)
    best = 0
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        _, fx, st = fe.find_minimum(x, max_iter=50)
        torch.cuda.synchronize(); w = time.perf_counter() - t0
        best = max(best, st["nit"] / w)
print(f"SCG_FUSED={os.environ.get('MFVGP_SCG_FUSED','1')} PDL={os.environ.get('MFVGP_PDL','1')}: {best:.0f} it/s ({1e6/best:.1f} us/it), nit={st['nit']} evals={st['func_eval']} fx={fx:.12g}")
