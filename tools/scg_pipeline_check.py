"""A/B of the pipelined SCG chains (MFVGP_SCG_PIPELINE=0/1): identical results, iterations/s."""
import contextlib, io, os, subprocess, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))

def run():
    import numpy as np, torch
    from helpers import make_free_energy, synthetic_l96, load_golden
    out = {}
    for name, D, M, iters in (("L96-500", 500, 16, 50), ("L96-40", 40, 20, 200), ("L96-2000", 2000, 64, 50)):
        g = synthetic_l96(D, M, seed=8192)
        fe = make_free_energy(g)
        x = torch.tensor(g["x0"], dtype=torch.float64, device="cuda")
        with contextlib.redirect_stdout(io.StringIO()):
            fe.find_minimum(x, max_iter=10)
            best = 0
            for _ in range(3):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                xo, fx, st = fe.find_minimum(x, max_iter=iters)
                torch.cuda.synchronize(); w = time.perf_counter() - t0
                best = max(best, st["nit"] / w)
        print(f"{name}: {best:.0f} it/s ({1e6/best:.1f} us/it) nit={st['nit']} evals={st['func_eval']} fx={fx!r} "
              f"xsum={float(xo.sum())!r}")
    for name in ("DW", "L63_T5", "OU", "L96_8"):
        g = load_golden("scg", name)
        fe = make_free_energy(g)
        with contextlib.redirect_stdout(io.StringIO()):
            fe.find_minimum(g["x0"], max_iter=int(g["max_iter"]))
            t0 = time.perf_counter()
            xo, fx, st = fe.find_minimum(g["x0"], max_iter=int(g["max_iter"]))
            w = time.perf_counter() - t0
        print(f"{name}: {w*1e3:.2f} ms nit={st['nit']} evals={st['func_eval']} fx={fx!r} xsum={float(np.sum(xo))!r}")

if __name__ == "__main__":
    if len(sys.argv) > 1:
        run()
    else:
        for pipe, cl in (("0", "0"), ("1", "0"), ("0", "1"), ("1", "1")):
            print(f"--- MFVGP_SCG_PIPELINE={pipe} MFVGP_SCG_CLUSTER={cl}")
            sys.stdout.flush()
            subprocess.run([sys.executable, __file__, "run"],
                           env=dict(os.environ, MFVGP_SCG_PIPELINE=pipe, MFVGP_SCG_CLUSTER=cl), check=False)
