"""Soak: many small random problems, SCG with pipelining off / default / forced and with grid /
cluster reductions -- pipelining must be bit-identical, all variants must take the same branches."""
import contextlib, io, os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from helpers import make_free_energy, synthetic_l96, load_golden  # noqa: E402

def run(g, iters):
    out = {}
    for cl in ("0", "1"):
        for pipe in ("0", "1", "2"):
            os.environ["MFVGP_SCG_PIPELINE"] = pipe
            os.environ["MFVGP_SCG_CLUSTER"] = cl
            fe = make_free_energy(g)
            with contextlib.redirect_stdout(io.StringIO()):
                x, fx, st = fe.find_minimum(g["x0"].copy(), max_iter=iters)
            out[cl, pipe] = (x, fx, int(st["nit"]), int(st["func_eval"]))
    bad = []
    for cl in ("0", "1"):
        a = out[cl, "0"]
        for pipe in ("1", "2"):
            b = out[cl, pipe]
            if not (np.array_equal(a[0], b[0]) and a[1] == b[1] and a[2:] == b[2:]):
                bad.append(("pipe", cl, pipe, a[1:], b[1:]))
    if out["0", "0"][2:] != out["1", "0"][2:]:
        bad.append(("branches", out["0", "0"][1:], out["1", "0"][1:]))
    return bad, out["1", "1"][1:]

rng = np.random.default_rng(2026)
nbad = 0
for k in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
    D = int(rng.integers(4, 70)); M = int(rng.integers(3, 14))
    g = synthetic_l96(D, M, seed=int(rng.integers(1 << 30)), partial=bool(rng.integers(2)))
    iters = int(rng.integers(10, 80))
    try:
        bad, res = run(g, iters)
    except RuntimeError as e:
        print(k, D, M, "raised", str(e)[:60]); continue
    print(k, f"D={D} M={M} iters={iters}", res, "OK" if not bad else bad, flush=True)
    nbad += bool(bad)
for name in ("DW", "OU", "L63_T5"):
    g = load_golden("scg", name)
    bad, res = run(g, int(g["max_iter"]))
    print(name, res, "OK" if not bad else bad, flush=True)
    nbad += bool(bad)
print("mismatches:", nbad)
