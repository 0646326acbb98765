"""ncu / trace target: one device SCG run of 12 iterations on L96 D=500 (M=16)."""
import contextlib, io, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from helpers import make_free_energy, synthetic_l96  # noqa: E402
g = synthetic_l96(500, 16, seed=8192)
fe = make_free_energy(g)
x = torch.tensor(g["x0"], dtype=torch.float64, device="cuda")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 12
with contextlib.redirect_stdout(io.StringIO()):
    for _ in range(2):
        _, fx, st = fe.find_minimum(x, max_iter=iters)
torch.cuda.synchronize()
print("nit", st["nit"], "fx", fx)
