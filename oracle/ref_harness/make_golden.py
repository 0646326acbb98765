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


# -- SCG starts that take the rare branches of scaled_cg.py ---------------------------------
def _truncate(setup, M):
    """the setup cut down to its first M observation times (x0 columns with it)"""
    sde = setup["sde"]
    D = int(np.atleast_1d(sde.sigma).size)
    M0 = len(setup["obs_times"])
    mp = setup["x0"][:D * (3 * M0 + 4)].reshape(D, 3 * M0 + 4)[:, :3 * M + 4]
    sp = setup["x0"][D * (3 * M0 + 4):].reshape(D, 2 * M0 + 3)[:, :2 * M + 3]
    s = dict(setup)
    s["obs_times"] = np.asarray(setup["obs_times"])[:M]
    s["obs_values"] = np.asarray(setup["obs_values"])[:M]
    s["x0"] = np.concatenate((mp.ravel(), sp.ravel()))
    return s


def scg_branch_case(setup, M, seed, amps_m, amps_s, max_iter, name):
    """
    SCG from a perturbed start (x0 + am N(0,1) on the means, + as N(0,1) on the log-variances;
    am, as drawn from amps_m, amps_s by default_rng(seed) first: the recipe of
    tools/scg_branch_search.py, which found these starts).  The reference's SCG.minimize runs
    unmodified and keeps no branch counters, so the counters stored beside its statistics come
    from the oracle's restatement (oracle/mfvgp_oracle.py scg_minimize), after checking here
    that the restatement follows the reference's run: same nit, func_eval and fx / dfx history.
    """
    from oracle import mfvgp_oracle as O
    harness.install()
    s = _truncate(setup, M)
    sde = s["sde"]
    D = int(np.atleast_1d(sde.sigma).size)
    nm = D * (3 * M + 4)
    r = np.random.default_rng(seed)
    am, asg = float(r.choice(amps_m)), float(r.choice(amps_s))
    x0 = s["x0"].copy()
    x0[:nm] += am * r.standard_normal(nm)
    x0[nm:] += asg * r.standard_normal(x0.size - nm)
    s["x0"] = x0
    fe = setups.free_energy_of(s, n_jobs=1)
    t0 = time.perf_counter()
    x, fx, stats = fe.find_minimum(x0.copy(), max_iter=max_iter)
    wall = time.perf_counter() - t0
    out = _inputs(s)
    o = O.FreeEnergyOracle(out["system"], out["sigma"], out["theta"], out["mu0"], out["tau0"],
                           out["obs_times"], out["obs_values"], out["obs_noise"],
                           None if out["obs_mask"].size == 0 else list(out["obs_mask"]))
    ev = {}
    xo, fxo, so = o.find_minimum(x0.copy(), max_iter=max_iter, events=ev)
    nit = int(stats["nit"])
    assert so["nit"] == nit and so["func_eval"] == int(stats["func_eval"]), (so["nit"], nit)
    assert np.abs(so["fx"][:nit] - stats["fx"][:nit]).max() <= 1e-9 * np.abs(stats["fx"][:nit]).max()
    assert np.abs(so["dfx"][:nit] - stats["dfx"][:nit]).max() <= 1e-7 * np.abs(stats["dfx"][:nit]).max()
    out.update(max_iter=max_iter, x_opt=x, fx_opt=float(fx), nit=nit, fx_hist=stats["fx"],
               dfx_hist=stats["dfx"], func_eval=int(stats["func_eval"]), ref_wall_s=wall,
               n_reject=ev["reject"], n_delta_le0=ev["delta<=0"], n_mu_ge0=ev["mu>=0"],
               n_restart=ev["restart"], start_seed=seed, start_am=am, start_as=asg)
    np.savez_compressed(GOLDEN_DIR / f"scg_{name}.npz", **out)
    print(f"scg_{name}: fx={fx:.12g} nit={nit} evals={stats['func_eval']} branches={ev} ({wall:.1f}s)")


def scg_hilltop_case(M, seed, sig, c, amp, sv, rn, max_iter, name):
    """
    delta <= 0 (scaled_cg.py:233-236) needs negative curvature along the search direction.  On
    sane inputs that was only found where the double well is genuinely non-convex: the mean
    path started on the hilltop between the wells (m = c ~ 0, the drift's unstable point), small
    uniform variances, weak observations (obs_noise rn) and system noise sig.  Found by a random
    search over (sig, c, amp, sv, rn) with the oracle; the recipe below rebuilds the start.
    """
    harness.install()
    C = harness.reference_classes()
    s = _truncate(setups.make_dw(), M)
    sde = C["DoubleWell"](sig, 1.0, r_seed=575)
    sde.make_trajectory(0.0, 10.0, 0.001)              # only sets the time window FreeEnergy checks
    s["sde"], s["obs_noise"] = sde, rn
    r = np.random.default_rng(seed)
    for grid in ([0.02, 0.05, 0.2, 0.8], [0.0, 0.2, 0.4, 0.58, 0.8], [0.02, 0.1], [0.01, 0.05, 0.2],
                 [0.04, 1.0, 25.0]):
        r.choice(grid)                                  # the search drew its parameters first
    nm = 3 * M + 4
    x0 = s["x0"].copy()
    x0[:nm] = c + amp * r.standard_normal(nm)
    x0[nm:] = np.log(sv) + 0.05 * r.standard_normal(x0.size - nm)
    s["x0"] = x0
    from oracle import mfvgp_oracle as O
    fe = setups.free_energy_of(s, n_jobs=1)
    t0 = time.perf_counter()
    x, fx, stats = fe.find_minimum(x0.copy(), max_iter=max_iter)
    wall = time.perf_counter() - t0
    out = _inputs(s)
    o = O.FreeEnergyOracle(out["system"], out["sigma"], out["theta"], out["mu0"], out["tau0"],
                           out["obs_times"], out["obs_values"], out["obs_noise"], None)
    ev = {}
    xo, fxo, so = o.find_minimum(x0.copy(), max_iter=max_iter, events=ev)
    nit = int(stats["nit"])
    assert so["nit"] == nit and so["func_eval"] == int(stats["func_eval"]), (so["nit"], nit, so["func_eval"], stats["func_eval"])
    assert np.abs(so["fx"][:nit] - stats["fx"][:nit]).max() <= 1e-9 * np.abs(stats["fx"][:nit]).max()
    assert ev["delta<=0"] >= 1
    out.update(max_iter=max_iter, x_opt=x, fx_opt=float(fx), nit=nit, fx_hist=stats["fx"],
               dfx_hist=stats["dfx"], func_eval=int(stats["func_eval"]), ref_wall_s=wall,
               n_reject=ev["reject"], n_delta_le0=ev["delta<=0"], n_mu_ge0=ev["mu>=0"],
               n_restart=ev["restart"], start_seed=seed)
    np.savez_compressed(GOLDEN_DIR / f"scg_{name}.npz", **out)
    print(f"scg_{name}: fx={fx:.12g} nit={nit} evals={stats['func_eval']} branches={ev} ({wall:.1f}s)")


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
    # rejected trial points (scaled_cg.py:254-279), mu >= 0 (:194-197), delta <= 0 (:233-236),
    # restart after dim_x successes (:361-364), convergence exit (:306-318)
    "scg_branch_DW3": lambda: scg_branch_case(
        setups.make_dw(), 3, 672241289, [0.2, 0.5, 1.0, 1.5, 3.0], [0.0, 0.3, 0.6, 1.0, 1.5], 80,
        "branch_DW3"),
    "scg_branch_DW4": lambda: scg_branch_case(
        setups.make_dw(), 4, 819696831, [0.2, 0.5, 1.0, 2.0, 3.0], [0.0, 0.2, 0.4, 0.6], 60,
        "branch_DW4"),
    "scg_branch_OU2": lambda: scg_branch_case(
        setups.make_ou(), 2, 630436052, [0.0, 0.5, 1.5], [0.0, 0.5, 1.0, 2.0], 60, "branch_OU2"),
    "scg_branch_L96_8": lambda: scg_branch_case(
        setups.make_l96(D=8, tf=2.0), 2, 335045, [0.2, 0.5, 1.0, 2.0], [0.6, 0.8, 1.0, 1.2], 40,
        "branch_L96_8"),
    "scg_hilltop_DW3": lambda: scg_hilltop_case(3, 38744149, 0.2, 0.0, 0.02, 0.05, 25.0, 25,
                                                "hilltop_DW3"),
    "scg_hilltop_DW3b": lambda: scg_hilltop_case(3, 899221749, 0.02, 0.0, 0.02, 0.01, 1.0, 25,
                                                 "hilltop_DW3b"),
}

if __name__ == "__main__":
    GOLDEN_DIR.mkdir(parents=True, exist_ok=True)
    for key in (sys.argv[1:] or list(CASES)):
        CASES[key]()
