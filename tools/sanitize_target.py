"""compute-sanitizer target: small evaluations through every kernel family."""
import contextlib, io, os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from helpers import load_golden, make_free_energy, synthetic_l96
import meanfieldvargp_b200 as mf
which = sys.argv[1] if len(sys.argv) > 1 else "all"
def run(g, x=None, scg=False):
    fe = make_free_energy(g)
    F, grad = fe.E_cost(g["x0"] if x is None else x)
    fe.E_cost(g["x0"] if x is None else x, output_gradients=False)
    if scg:
        with contextlib.redirect_stdout(io.StringIO()):
            fe.find_minimum(g["x0"], max_iter=10)
    return F
if which in ("all", "coop"):
    print("coop", run(synthetic_l96(67, 6, seed=1), scg=True))
    g = load_golden("eval", "L96_6_refine"); print("coop refine", run(g))
if which in ("all", "walk"):
    os.environ["MFVGP_PATH"] = "walk"; os.environ["MFVGP_FTILE"] = "64,3"
    print("walk", run(synthetic_l96(130, 11, seed=2)))
    g = synthetic_l96(67, 12, seed=79); x = g["x0"].copy()
    sp = x[67 * 40:].reshape(67, 27); sp[:, [3, 11]] = np.log(0.03); print("walk refine", run(g, x))
    os.environ.pop("MFVGP_PATH"); os.environ.pop("MFVGP_FTILE")
if which in ("all", "small"):
    for name in ("DW", "L63", "DW_refine"):
        print(name, run(load_golden("eval", name)))
if which in ("all", "split"):
    os.environ["MFVGP_PATH"] = "split"; os.environ["MFVGP_COOP"] = "0"
    print("split", run(synthetic_l96(67, 6, seed=1)))
    os.environ.pop("MFVGP_PATH"); os.environ.pop("MFVGP_COOP")
if which in ("all", "next"):
    tk, x = mf.simulate_l96(10.0, 8.0, 33, 0.0, 0.05, 0.002, r_seed=1)
    mp = np.random.default_rng(0).normal(size=(5, 3 * 4 + 4)); sp = 4 + np.random.default_rng(1).random((5, 2 * 4 + 3))
    print("next", float(x[-1, 0]), mf.reconstruct(mp, sp, 0.0, [0.2, 0.4, 0.6, 0.8], 1.0, num=7)[1].shape)
