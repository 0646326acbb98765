"""
TEST INFRASTRUCTURE (oracle side): golden vectors for the theta / Sigma gradients
(SURVEY.md 8(f) rank 4), generated from the REFERENCE's own material.  Run here only:

    python -m oracle.ref_harness.make_golden_params [case ...]

The reference implements no hyper-parameter gradient (free_energy.py:241-279, :745-748 are the
hooks), so there is no reference output to record.  Two independent pins instead:

  1. ``sympy``: the reference's own energy expressions (energy_functions/*.sym, decoded where
     they lie by ``symdecode``) are called with SymPy symbols, differentiated with respect to
     the theta and Sigma arguments of their signature (stochastic_process.py:312-326),
     lambdified, and integrated per interval with ``scipy.integrate.quad_vec`` with the
     arguments of free_energy.py:405-417; assembled with the 1/Sigma contraction and the 1/2 of
     free_energy.py:423-458:  dF/dtheta_k, dF/dSigma_j.
  2. ``fd``: central differences of the UNMODIFIED reference's ``FreeEnergy.E_cost(x,
     output_gradients=False)`` through its ``sde_theta`` / ``sde_sigma`` setters.

Both are stored in tests/golden/params_<case>.npz beside the inputs.
"""
import sys
import time
from pathlib import Path

import numpy as np
import sympy as sp
from scipy.integrate import quad_vec

from . import harness, setups, symdecode
from .make_golden import GOLDEN_DIR, SYS_OF, _inputs

ENERGY_DIR = harness.REFERENCE_ROOT / "src" / "dynamical_systems" / "energy_functions"
FILES = {"DW": ["DW_Esde_0.sym"], "OU": ["OU_Esde_0.sym"],
         "L63": ["L63_Esde_0.sym", "L63_Esde_1.sym", "L63_Esde_2.sym"],
         "L96": ["L96_Esde_0.sym"]}
KDIMS = {"DW": 1, "OU": 1, "L63": 3, "L96": 4}          # state dimensions in one expression
NTHETA = {"DW": 1, "OU": 2, "L63": 3, "L96": 1}


def _derivative_functions(system):
    """For every energy expression: lambdified [E, dE/dtheta_k..., dE/dSig_l...] of (t, *args)."""
    k, nth = KDIMS[system], NTHETA[system]
    names = (["t", "h0", "h1", "h2", "h3", "c0", "c1", "c2"]
             + [f"d{j}m{i}" for j in range(k) for i in range(4)]
             + [f"d{j}s{i}" for j in range(k) for i in range(3)]
             + [f"Sig{j}" for j in range(k)] + [f"th{i}" for i in range(nth)])
    syms = sp.symbols(names)
    sig_syms = syms[8 + 7 * k: 8 + 8 * k]
    th_syms = syms[8 + 8 * k:]
    out = []
    for fname in FILES[system]:
        with open(ENERGY_DIR / fname, "rb") as fh:
            func = symdecode.load(fh)
        expr = func(*syms)
        exprs = [expr] + [sp.diff(expr, v) for v in th_syms] + [sp.diff(expr, v) for v in sig_syms]
        out.append(sp.lambdify(syms, exprs, "numpy", cse=True))
    return out


def _sympy_gradients(setup):
    sde = setup["sde"]
    system = SYS_OF[type(sde).__name__]
    k, nth = KDIMS[system], NTHETA[system]
    funcs = _derivative_functions(system)
    sigma = np.atleast_1d(sde.sigma).astype(float)
    theta = np.atleast_1d(sde.theta).astype(float)
    D = sigma.size
    obs_times = np.asarray(setup["obs_times"], dtype=float)
    M = obs_times.size
    x = setup["x0"]
    mean_pts = x[:D * (3 * M + 4)].reshape(D, 3 * M + 4)
    vars_pts = np.exp(x[D * (3 * M + 4):].reshape(D, 2 * M + 3))
    inv_sigma = 1.0 / sigma
    dth, dsig, Esde = np.zeros(nth), np.zeros(D), 0.0
    neval = []
    kw = dict(limit=100, epsabs=1.0e-06, epsrel=1.0e-06, full_output=True)
    for n in range(M - 1):                                   # free_energy.py:490
        ti, tj = obs_times[n], obs_times[n + 1]
        delta_t = np.abs(tj - ti)                            # free_energy.py:386-398
        h, c = float(delta_t / 3.0), float(delta_t / 2.0)
        tvars = (ti, ti + h, ti + 2 * h, tj, ti, ti + c, tj)
        mp, spn = mean_pts[:, 3 * n:3 * n + 4], vars_pts[:, 2 * n:2 * n + 3]

        def dims_of(i):
            if system == "L96":                              # lorenz_96.py:232-235
                return [i, (i + 1) % D, (i - 1) % D, (i - 2) % D]
            return list(range(D))

        def family(which):
            """which = 0: E_i [D]; 1: dE_i/dtheta_k [D, nth]; 2: dE_i/dSig_l [D, k]"""
            def f(t):
                rows = []
                for i in range(D):
                    idx = dims_of(i)
                    fn = funcs[0] if system == "L96" else funcs[i]
                    vals = fn(t, *tvars, *mp[idx].ravel(), *spn[idx].ravel(), *sigma[idx], *theta)
                    vals = [np.broadcast_to(np.asarray(v, dtype=float), np.shape(t)) for v in vals]
                    if which == 0:
                        rows.append(vals[0])
                    elif which == 1:
                        rows.append(np.stack(vals[1:1 + nth]))
                    else:
                        rows.append(np.stack(vals[1 + nth:]))
                return np.array(rows)
            return f

        r = quad_vec(family(0), ti, tj, **kw)
        i_E, ne = r[0], [r[2].neval]
        r = quad_vec(family(1), ti, tj, **kw)
        i_th = r[0]
        ne.append(r[2].neval)
        r = quad_vec(family(2), ti, tj, **kw)
        i_sg = r[0]
        ne.append(r[2].neval)
        neval.append(ne)
        Esde += 0.5 * float(inv_sigma @ i_E)
        dth += 0.5 * (inv_sigma @ i_th)
        dsig += -0.5 * inv_sigma ** 2 * i_E
        for i in range(D):
            for l, j in enumerate(dims_of(i)):
                dsig[j] += 0.5 * inv_sigma[i] * i_sg[i, l]
    return dth, dsig, Esde, np.array(neval)


def _fd_gradients(setup, n_sigma=6, rel_step=1.0e-5, n_jobs=None):
    """central differences of the unmodified reference's E_cost through its setters"""
    fe = setups.free_energy_of(setup, n_jobs=n_jobs)
    x = setup["x0"]
    theta0 = np.array(np.atleast_1d(fe.sde_theta), dtype=float)
    sigma0 = np.array(np.atleast_1d(fe.sde_sigma), dtype=float)

    def F():
        return float(fe.E_cost(x, output_gradients=False))

    fd_th = np.zeros(theta0.size)
    for k in range(theta0.size):
        h = rel_step * max(abs(theta0[k]), 1.0)
        tp, tm = theta0.copy(), theta0.copy()
        tp[k] += h
        tm[k] -= h
        fe.sde_theta = tp          # 1-d arrays: single_interval unpacks *theta (:401)
        Fp = F()
        fe.sde_theta = tm
        Fm = F()
        fd_th[k] = (Fp - Fm) / (2 * h)
    fe.sde_theta = theta0
    D = sigma0.size
    idx = np.unique(np.linspace(0, D - 1, min(D, n_sigma)).astype(int))
    fd_sg = np.zeros(idx.size)
    for a, j in enumerate(idx):
        h = rel_step * sigma0[j]
        sp_, sm_ = sigma0.copy(), sigma0.copy()
        sp_[j] += h
        sm_[j] -= h
        fe.sde_sigma = sp_
        Fp = F()
        fe.sde_sigma = sm_
        Fm = F()
        fd_sg[a] = (Fp - Fm) / (2 * h)
    fe.sde_sigma = sigma0
    return fd_th, idx, fd_sg


def param_case(setup, name=None, n_jobs=None):
    harness.install()
    t0 = time.perf_counter()
    dth, dsig, Esde, neval = _sympy_gradients(setup)
    fe = setups.free_energy_of(setup, n_jobs=n_jobs)
    D, M = fe.dim_D, fe.num_M
    x = setup["x0"]
    mp = x[:fe.num_mp].reshape(D, 3 * M + 4)
    sv = np.exp(x[fe.num_mp:].reshape(D, 2 * M + 3))
    Esde_ref = float(fe.E_sde(mp, sv, output_gradients=False)[0])
    # the SymPy route integrates the reference's expressions exactly as the reference does
    assert abs(Esde - Esde_ref) <= 1e-12 * abs(Esde_ref), (Esde, Esde_ref)
    fd_th, fd_idx, fd_sg = _fd_gradients(setup, n_jobs=n_jobs)
    out = _inputs(setup)
    out.update(dF_dtheta=dth, dF_dsigma=dsig, Esde=Esde_ref, neval=neval, fd_dtheta=fd_th,
               fd_sigma_idx=fd_idx, fd_dsigma=fd_sg)
    name = name or setup["name"]
    np.savez_compressed(GOLDEN_DIR / f"params_{name}.npz", **out)
    e_th = np.abs(dth - fd_th).max() / max(np.abs(fd_th).max(), 1e-300)
    e_sg = np.abs(dsig[fd_idx] - fd_sg).max() / max(np.abs(fd_sg).max(), 1e-300)
    print(f"params_{name}: dF/dtheta={dth} (fd rel {e_th:.1e}) |dF/dSigma|max={np.abs(dsig).max():.6g} "
          f"(fd rel {e_sg:.1e}) neval={sorted(set(neval.ravel()))} ({time.perf_counter() - t0:.1f}s)")


def _refine_dw():
    from .make_golden import refine_dw
    return refine_dw()


def _refine_l96():
    from .make_golden import refine_l96
    return refine_l96()


CASES = {
    "DW": lambda: param_case(setups.make_dw()),
    "OU": lambda: param_case(setups.make_ou()),
    "L63": lambda: param_case(setups.make_l63(tf=5.0), name="L63_T5"),
    "L96_8": lambda: param_case(setups.make_l96(D=8, tf=2.0)),
    "L96_40": lambda: param_case(setups.make_l96(D=40, tf=5.0), n_jobs=8),
    "L96_67p": lambda: param_case(setups.make_l96(D=67, tf=1.5, partial=True), n_jobs=8),
    "DW_refine": lambda: param_case(_refine_dw()),
    "L96_6_refine": lambda: param_case(_refine_l96()),
}

if __name__ == "__main__":
    GOLDEN_DIR.mkdir(parents=True, exist_ok=True)
    for key in (sys.argv[1:] or list(CASES)):
        CASES[key]()
