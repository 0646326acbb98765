"""
TEST INFRASTRUCTURE (oracle side): generates tests/golden/*.npz by running the
UNMODIFIED reference (through ``harness``) on seeded inputs.  Run here only:

    python -m oracle.ref_harness.make_golden [case ...]

Each file holds the inputs (so the GPU box, which has no /root/reference, can
replay them) and the reference's outputs: ``E_cost`` value and gradient,
``E_sde`` pieces, ``E_kl0``/``E_obs`` values, ``quad_vec`` evaluation counts
per (interval, integrand), and for the ``scg_*`` cases the full SCG statistics.
"""
import sys
import time
from pathlib import Path

import numpy as np

from . import harness, setups

GOLDEN_DIR = Path(__file__).resolve().parents[2] / "tests" / "golden"
SYS_OF = {"DoubleWell": "DW", "OrnsteinUhlenbeck": "OU", "Lorenz63": "L63",
          "Lorenz96": "L96"}


class _NevalRecorder(object):
    """Wraps scipy's quad_vec inside the reference module to log ``neval``."""

    def __init__(self):
        import src.var_bayes.free_energy as fe_mod
        from scipy.integrate import quad_vec
        self.fe_mod, self.orig, self.log = fe_mod, quad_vec, []

        def wrapped(f, a, b, **kw):
            res = self.orig(f, a, b, full_output=True, **kw)
            self.log.append(res[2].neval)
            return res[0], res[1]
        fe_mod.quad_vec = wrapped

    def close(self):
        self.fe_mod.quad_vec = self.orig


def _inputs(setup):
    sde = setup["sde"]
    mask = setup["obs_mask"]
    return dict(system=SYS_OF[type(sde).__name__],
                sigma=np.atleast_1d(sde.sigma).astype(float),
                theta=np.atleast_1d(sde.theta).astype(float),
                mu0=np.asarray(setup["mu0"], dtype=float),
                tau0=np.asarray(setup["tau0"], dtype=float),
                obs_times=np.asarray(setup["obs_times"], dtype=float),
                obs_values=np.asarray(setup["obs_values"], dtype=float),
                obs_noise=np.asarray(setup["obs_noise"], dtype=float),
                obs_mask=np.asarray([] if mask is None else mask, dtype=bool),
                x0=np.asarray(setup["x0"], dtype=float))


def eval_case(setup, name=None):
    """One E_cost evaluation with all intermediate pieces."""
    harness.install()
    fe = setups.free_energy_of(setup, n_jobs=1)
    rec = _NevalRecorder()
    try:
        x = setup["x0"]
        t0 = time.perf_counter()
        F, grad = fe.E_cost(x)
        wall = time.perf_counter() - t0
        neval = np.array(rec.log, dtype=np.int64).reshape(-1, 3)
        D, M = fe.dim_D, fe.num_M
        mp = x[:fe.num_mp].reshape(D, 3 * M + 4)
        sp = np.exp(x[fe.num_mp:].reshape(D, 2 * M + 3))
        Esde, dm, dsv = fe.E_sde(mp, sp)
        E0, _, _ = fe.E_kl0(mp[:, 0], sp[:, 0])
        Eobs, _, _ = fe.E_obs(mp[fe.ix_ikm], sp[fe.ix_iks])
        F_only = fe.E_cost(x, output_gradients=False)
    finally:
        rec.close()
    out = _inputs(setup)
    out.update(F=float(F), grad=grad, Esde=float(Esde), dEsde_dm=dm, dEsde_ds=dsv,
               E0=float(E0), Eobs=float(Eobs), F_only=float(F_only), neval=neval,
               ref_wall_s=wall)
    name = name or setup["name"]
    np.savez_compressed(GOLDEN_DIR / f"eval_{name}.npz", **out)
    print(f"eval_{name}: F={F:.12g} len(x)={x.size} neval={sorted(set(neval.ravel()))} "
          f"({wall:.2f}s)")


def scg_case(setup, max_iter, name=None):
    harness.install()
    fe = setups.free_energy_of(setup, n_jobs=1)
    t0 = time.perf_counter()
    x, fx, stats = fe.find_minimum(setup["x0"].copy(), max_iter=max_iter)
    wall = time.perf_counter() - t0
    out = _inputs(setup)
    out.update(max_iter=max_iter, x_opt=x, fx_opt=float(fx), nit=int(stats["nit"]),
               fx_hist=stats["fx"], dfx_hist=stats["dfx"],
               func_eval=int(stats["func_eval"]), ref_wall_s=wall)
    name = name or setup["name"]
    np.savez_compressed(GOLDEN_DIR / f"scg_{name}.npz", **out)
    print(f"scg_{name}: fx={fx:.12g} nit={stats['nit']} evals={stats['func_eval']} ({wall:.1f}s)")


# -- special inputs -----------------------------------------------------------
def _with_x(setup, x, name):
    s = dict(setup)
    s["x0"], s["name"] = x, name
    return s


def refine_dw():
    """Variance support points that differ by >10x: quad_vec refines (a-4)."""
    s = setups.make_dw()
    M = len(s["obs_times"])
    n_mp = 3 * M + 4
    x = s["x0"].copy()
    sp = np.full(2 * M + 3, 0.3)
    sp[1::4] = 1.0e-2
    sp[7] = 1.0e-3
    x[n_mp:] = np.log(sp)
    return _with_x(s, x, "DW_refine")


def refine_l96():
    s = setups.make_l96(D=6, tf=2.0, sigma=0.5)
    D, M = 6, len(s["obs_times"])
    n_mp = D * (3 * M + 4)
    rng = np.random.default_rng(7)
    sp = 0.3 * np.ones((D, 2 * M + 3))
    sp[:, 1::2] = 10.0 ** rng.uniform(-2.5, -0.5, size=sp[:, 1::2].shape)
    x = s["x0"].copy()
    x[n_mp:] = np.log(sp).ravel()
    return _with_x(s, x, "L96_6_refine")


def irregular_l96():
    """Non-uniform observation spacing: the path only sees Delta_n (a-7)."""
    s = setups.make_l96(D=5, tf=2.0)
    rng = np.random.default_rng(3)
    t = np.cumsum(rng.uniform(0.08, 0.45, size=len(s["obs_times"])))
    s["obs_times"] = t
    s["name"] = "L96_5_irregular"
    return s


CASES = {
    "DW": lambda: eval_case(setups.make_dw()),
    "OU": lambda: eval_case(setups.make_ou()),
    "L63": lambda: eval_case(setups.make_l63()),
    "L96_4": lambda: eval_case(setups.make_l96(D=4, tf=1.0)),
    "L96_8": lambda: eval_case(setups.make_l96(D=8, tf=2.0)),
    "L96_40": lambda: eval_case(setups.make_l96(D=40, tf=5.0)),
    "L96_67p": lambda: eval_case(setups.make_l96(D=67, tf=1.5, partial=True)),
    "L96_500p": lambda: eval_case(setups.make_l96(D=500, tf=4.0, partial=True)),
    "DW_refine": lambda: eval_case(refine_dw()),
    "L96_6_refine": lambda: eval_case(refine_l96()),
    "L96_5_irregular": lambda: eval_case(irregular_l96()),
    "scg_DW": lambda: scg_case(setups.make_dw(), 40),
    "scg_OU": lambda: scg_case(setups.make_ou(), 25),
    "scg_L63": lambda: scg_case(setups.make_l63(tf=5.0), 15, name="L63_T5"),
    "scg_L96_8": lambda: scg_case(setups.make_l96(D=8, tf=2.0), 15),
}

if __name__ == "__main__":
    GOLDEN_DIR.mkdir(parents=True, exist_ok=True)
    for key in (sys.argv[1:] or list(CASES)):
        CASES[key]()
