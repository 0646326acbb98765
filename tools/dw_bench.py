"""DW / L63 golden problems: E_cost latency and SCG wall time."""
import contextlib, io, os, sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from helpers import load_golden, make_free_energy
for name in ("DW", "L63_T5"):
    g = load_golden("scg", name)
    fe = make_free_energy(g)
    x = torch.tensor(g["x0"], dtype=torch.float64, device="cuda")
    for _ in range(10): fe.E_cost(x)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(300): fe.E_cost(x)
    torch.cuda.synchronize(); ev = (time.perf_counter() - t0) / 300 * 1e6
    with contextlib.redirect_stdout(io.StringIO()):
        fe.find_minimum(x, max_iter=20)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        _, fx, st = fe.find_minimum(x, max_iter=150)
        torch.cuda.synchronize(); w = time.perf_counter() - t0
    l0 = fe.kernel_launches; fe.E_cost(x); l1 = fe.kernel_launches
    print(f"{name} COOP={os.environ.get('MFVGP_COOP','1')}: E_cost {ev:.1f} us back to back, {l1-l0} launches; SCG {st['nit']} its ({st['func_eval']} evals) in {w*1e3:.1f} ms = {st['nit']/w:.0f} it/s, fx={fx:.6g}")
