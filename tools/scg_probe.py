"""How often does the guard flag intervals along an SCG run, and what does it cost?"""
import contextlib, io, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from helpers import make_free_energy, make_oracle, synthetic_l96, load_golden

def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t)/reps*1e6

for label, g in (("synthetic500", synthetic_l96(500, 16, seed=8192)), ("golden500", load_golden("eval","L96_500p"))):
    fe = make_free_energy(g)
    x = torch.tensor(g["x0"], dtype=torch.float64, device="cuda")
    for its in (5, 20, 50):
        with contextlib.redirect_stdout(io.StringIO()):
            xo, fx, st = fe.find_minimum(x, max_iter=its)
        fe.E_cost(xo); torch.cuda.synchronize()
        info = fe.refine_info()
        flagged = (info > 0).sum(axis=0)
        t_x0 = timeit(lambda: fe.E_cost(x)); t_xo = timeit(lambda: fe.E_cost(xo))
        sp = torch.exp(xo[fe.num_mp:]).reshape(500, -1)
        print(f"{label} its={its} fx={fx:.6g} flagged(E,dS)={flagged.tolist()} of {info.shape[0]} bisections max={info.max()} "
              f"E_cost us: at x0 {t_x0:.1f}, at x_opt {t_xo:.1f}; var range [{sp.min().item():.3g},{sp.max().item():.3g}]")
        if its == 50 and label == "synthetic500":
            o = make_oracle(g); o.E_cost(xo.cpu().numpy()); nev = np.array(o.neval)
            print("   scipy neval histogram E:", np.unique(nev[:,0], return_counts=True), " dS:", np.unique(nev[:,2], return_counts=True))
