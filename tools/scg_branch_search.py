"""
Searches (with the numpy oracle, CPU) for SCG starts on which the reference's SCG.minimize takes
its rare branches -- rejected trial point (scaled_cg.py:254-279), delta <= 0 (:233-236), mu >= 0
(:194-197), restart after dim_x successes (:361-364) -- so that they can be pinned as reference
goldens (oracle/ref_harness/make_golden.py scg_branch_*).

    python tools/scg_branch_search.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import load_golden, make_oracle  # noqa: E402


def scg_branches(func, x0, max_it, x_tol=1e-5, f_tol=1e-5):
    """branch counters of the oracle's SCG restatement (oracle/mfvgp_oracle.py scg_minimize)"""
    from oracle.mfvgp_oracle import scg_minimize
    ev = {}
    _, _, st = scg_minimize(func, x0, max_it, x_tol, f_tol, events=ev)
    ev["nit"], ev["evals"] = int(st["nit"]), int(st["func_eval"])
    return ev


def truncated(g, M):
    """the golden problem cut down to its first M observation times"""
    D = int(np.atleast_1d(g["sigma"]).size)
    M0 = int(g["obs_times"].size)
    Nm0, Ns0 = 3 * M0 + 4, 2 * M0 + 3
    mp = g["x0"][:D * Nm0].reshape(D, Nm0)[:, :3 * M + 4]
    sp = g["x0"][D * Nm0:].reshape(D, Ns0)[:, :2 * M + 3]
    g2 = dict(g)
    g2["obs_times"] = g["obs_times"][:M]
    g2["obs_values"] = g["obs_values"][:M]
    g2["x0"] = np.concatenate((mp.ravel(), sp.ravel()))
    return g2


def main():
    rng = np.random.default_rng(2027)
    for name, M, iters in (("DW", 2, 60), ("OU", 2, 60), ("DW", 4, 60), ("DW", 20, 60),
                           ("L63_T5", 3, 40), ("L96_8", 3, 40)):
        g0 = load_golden("scg", name)
        g = truncated(g0, M) if M < g0["obs_times"].size else g0
        D = int(np.atleast_1d(g["sigma"]).size)
        nm = D * (3 * int(g["obs_times"].size) + 4)
        for trial in range(24):
            seed = int(rng.integers(1 << 30))
            r = np.random.default_rng(seed)
            am, asg = float(r.choice([0.0, 0.5, 1.5])), float(r.choice([0.0, 0.5, 1.0, 2.0]))
            x0 = g["x0"].copy()
            x0[:nm] += am * r.standard_normal(nm)
            x0[nm:] += asg * r.standard_normal(x0.size - nm)
            o = make_oracle(g, adaptive=True)
            try:
                ev = scg_branches(o.E_cost, x0, iters)
            except (RuntimeError, FloatingPointError, ZeroDivisionError) as ex:
                continue
            tags = [k for k in ("reject", "delta<=0", "mu>=0", "restart") if ev[k]]
            if tags:
                print(f"{name} M={M} n={x0.size} seed={seed} am={am} as={asg}: {ev}", flush=True)


if __name__ == "__main__":
    main()
