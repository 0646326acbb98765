"""Looks for starts from which the device SCG rejects trial points (Delta < 0) and still finishes."""
import contextlib, io, os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from helpers import load_golden, make_free_energy  # noqa: E402
os.environ["MFVGP_SCG_PIPELINE"] = "0"
for name, dims in (("L63_T5", 3), ("DW", 1), ("L96_8", 8)):
    g = load_golden("scg", name)
    n = g["x0"].size
    nm = dims * (3 * int(g["obs_times"].size) + 4)
    for base in (0.5, 1.0, 1.5, 2.0, 2.5, 3.0):
        for seed in (1, 2, 3, 4):
            rng = np.random.default_rng(seed)
            g2 = dict(g); x0 = g["x0"].copy()
            x0[nm:] = x0[nm:] + base * rng.normal(0.0, 1.0, n - nm)
            g2["x0"] = x0
            fe = make_free_energy(g2)
            try:
                with contextlib.redirect_stdout(io.StringIO()):
                    x, fx, st = fe.find_minimum(x0.copy(), max_iter=40)
                fails = (3 * st["nit"] + 1 - st["func_eval"]) // 2
                if fails:
                    print(name, base, seed, "nit", st["nit"], "evals", st["func_eval"], "rejected", fails, "fx", fx)
            except RuntimeError as e:
                pass
