"""ncu / timing target: batched evaluation of L96 D=500 (M=16), B states per launch pair."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from helpers import make_free_energy, synthetic_l96  # noqa: E402
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
g = synthetic_l96(500, 16, seed=8192)
fe = make_free_energy(g)
rng = np.random.default_rng(11)
X = torch.as_tensor(g["x0"][None, :] + 0.05 * rng.standard_normal((B, g["x0"].size)), device="cuda")
for _ in range(3):
    F, G = fe.E_cost_batch(X, chunk=B)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    F, G = fe.E_cost_batch(X, chunk=B)
e1.record(); torch.cuda.synchronize()
print(f"B={B}: {e0.elapsed_time(e1) / 10:.3f} ms per batch, {B * 1e4 / e0.elapsed_time(e1):.0f} evals/s")
