"""e2e A/B: FreeEnergy.E_cost(numpy x) calls per second, pinned and pageable x (run with env switches).
Tried with it and NOT kept: the kernels reading a page-locked x straight over PCIe instead of the
H2D copy in front of them -- 80.5 us per call against 48.0 (8-byte strided reads over PCIe); and a
separate pull kernel (one 16-byte load per thread, all in flight at once) in place of the
copy-engine transfer -- 48.4 us against 46.4 (median of 15 blocks, twice each, same box)."""
import sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from meanfieldvargp_b200.synthetic import make_free_energy, synthetic_l96
g = synthetic_l96(500, 16, seed=8192)
fe = make_free_energy(g)
x_pin = torch.tensor(g["x0"], dtype=torch.float64).pin_memory().numpy()
x_pg = np.array(g["x0"])
F0, g0 = fe.E_cost(x_pg)
for name, xh in (("pinned", x_pin), ("pageable", x_pg)):
    for _ in range(20): fe.E_cost(xh)
    walls = []
    for _ in range(15):
        t0 = time.perf_counter()
        for _ in range(50): F, gr = fe.E_cost(xh)
        walls.append(time.perf_counter() - t0)
    walls.sort()
    assert F == F0 and np.array_equal(gr, g0)
    print(f"{name}: median {50/walls[7]:.0f} evals/s ({1e6*walls[7]/50:.1f} us), best {50/walls[0]:.0f}")
