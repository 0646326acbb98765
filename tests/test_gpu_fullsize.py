"""
Full-size checks (BASELINE.json configs[4]: Lorenz 96, D=8192, long horizon): the default
(walk) kernel against the neighbour-sparse oracle at D=8192 for short horizons (M = 9, 33: the
oracle takes seconds there), and at the long horizon through size-independent properties:

* the walk kernel (default) and the split kernel + assemble_kernel, two independent
  implementations of the same quadrature, agree to rounding;
* Lorenz 96 is circularly symmetric (lorenz_96.py:232-235): rolling the state dimensions
  (with sigma, the prior and the observations) leaves F unchanged and rolls the gradient;
* E_cost(x, output_gradients=False) returns the same F;
* the safe_log clamp of E_kl0 / E_obs at D=8192 (utilities.py:73-76).
"""
import numpy as np
import pytest

from helpers import make_free_energy, make_oracle, rel_err, synthetic_l96, worst_elem_err

pytestmark = pytest.mark.gpu

D, M = 8192, 513        # 4.2 M cells; the bench's M=4097 differs only in the number of runs


@pytest.fixture(scope="module")
def problem():
    return synthetic_l96(D, M, seed=8192)


def _eval(g, x, env=None, monkeypatch=None):
    if env:
        for k, v in env.items():
            monkeypatch.setenv(k, v)
    fe = make_free_energy(g)
    if env:
        for k in env:
            monkeypatch.delenv(k)
    F, grad = fe.E_cost(x)
    return fe, F, grad


def test_walk_equals_split_at_full_size(problem, monkeypatch):
    g = problem
    fe_w, Fw, gw = _eval(g, g["x0"])
    assert fe_w.describe()["path"] == "walk"
    fe_s, Fs, gs = _eval(g, g["x0"], {"MFVGP_PATH": "split"}, monkeypatch)
    assert fe_s.describe()["path"] == "split"
    assert abs(Fw - Fs) <= 1e-13 * abs(Fs)
    assert np.abs(gw - gs).max() <= 1e-12 * np.abs(gs).max()
    assert abs(fe_w.E_cost(g["x0"], output_gradients=False) - Fw) <= 1e-13 * abs(Fw)
    # columns the reference never integrates receive only observation terms (free_energy.py:711-712)
    Nm = 3 * M + 4
    gm = gw[:D * Nm].reshape(D, Nm)
    assert np.all(gm[:, [Nm - 6, Nm - 5, Nm - 3, Nm - 2, Nm - 1]] == 0.0)
    unobserved = np.array([not b for b in g["obs_mask"]])
    assert np.all(gm[unobserved, Nm - 4] == 0.0) and np.all(gm[~unobserved, Nm - 4] != 0.0)


def test_circular_shift_invariance(problem):
    g = problem
    k = 1237                                      # not a multiple of the dim tile (122)
    Nm, Ns = 3 * M + 4, 2 * M + 3
    mp = g["x0"][:D * Nm].reshape(D, Nm)
    sp = g["x0"][D * Nm:].reshape(D, Ns)
    mask = np.array(g["obs_mask"])
    y_full = np.zeros((M, D))
    y_full[:, mask] = g["obs_values"]
    g2 = dict(g)
    mask2 = np.roll(mask, k)
    g2["obs_mask"] = [bool(b) for b in mask2]
    g2["obs_values"] = np.roll(y_full, k, axis=1)[:, mask2]
    x2 = np.concatenate((np.roll(mp, k, axis=0).ravel(), np.roll(sp, k, axis=0).ravel()))
    _, F1, g1 = _eval(g, g["x0"])
    _, F2, gr2 = _eval(g2, x2)
    assert abs(F1 - F2) <= 1e-12 * abs(F1)
    g1m = np.roll(g1[:D * Nm].reshape(D, Nm), k, axis=0)
    g1s = np.roll(g1[D * Nm:].reshape(D, Ns), k, axis=0)
    ref = np.concatenate((g1m.ravel(), g1s.ravel()))
    assert np.abs(gr2 - ref).max() <= 1e-12 * np.abs(ref).max()


def test_safe_log_clamps_at_full_size(problem):
    """prod(tau0/s0) underflows and prod(obs_noise) overflows at this size; the reference
    returns the clamped logs (+-690.7755...), and so must we."""
    g = problem
    fe, F, _ = _eval(g, g["x0"])
    Nm, Ns = 3 * M + 4, 2 * M + 3
    mp = g["x0"][:D * Nm].reshape(D, Nm)
    s = np.exp(g["x0"][D * Nm:].reshape(D, Ns))
    clamp = np.log(1e300)
    # E_kl0, free_energy.py:309-310, with safe_log(prod(.)) = clip(sum log, -clamp, clamp)
    E0 = 0.5 * (np.clip(np.sum(np.log(g["tau0"] / s[:, 0])), -clamp, clamp)
                + np.sum(((mp[:, 0] - g["mu0"]) ** 2 + s[:, 0] - g["tau0"]) / g["tau0"]))
    assert np.sum(np.log(g["tau0"] / s[:, 0])) < -clamp          # the clamp is active
    assert abs(fe.last_terms["E0"] - E0) <= 1e-12 * abs(E0)
    # E_obs, free_energy.py:594-612
    mask = np.array(g["obs_mask"])
    d = int(mask.sum())
    ik, iks = 3 * np.arange(1, M + 1), 2 * np.arange(1, M + 1)
    R = g["obs_noise"]
    quad = np.sum(((g["obs_values"].T - mp[mask][:, ik]) ** 2 + s[mask][:, iks]) / R[:, None])
    assert d * np.log(2.0) > clamp                                # prod(R) overflows
    Eobs = 0.5 * (quad + M * (d * np.log(2.0 * np.pi) + clamp))
    assert abs(fe.last_terms["Eobs"] - Eobs) <= 1e-12 * abs(Eobs)


# ---- the walk kernel against the oracle at D=8192 (wrap-around halo over many dim tiles) --------
@pytest.mark.parametrize("M_,ftile,ntiles", [(9, None, 142), (33, None, 142), (9, "128,2", 68)])
def test_walk_kernel_equals_oracle_at_D8192(M_, ftile, ntiles, monkeypatch):
    """
    lorenz_96.py:221-236 (circular neighbours) over 142 tiles of 58 rows (the default at this D)
    and 68 tiles of 122 rows, free_energy.py:651-751.
    Two bands of rows get a steep variance mid-point so that scipy's quad_vec bisects further on
    those intervals (the oracle calls the real quad_vec): the guard must flag them and
    refine_kernel must reproduce the refined integrals.  Tolerance 1e-10 (north_star) on F, on the
    gradient in the max-norm and on its worst single element.
    """
    g = synthetic_l96(D, M_, seed=8192 + M_)
    Nm, Ns = 3 * M_ + 4, 2 * M_ + 3
    x = g["x0"].copy()
    sp = x[D * Nm:].reshape(D, Ns)
    sp[100:103, 5] = np.log(0.3)                  # interval 2: the dS integrand refines
    sp[4000:4002, 9] = np.log(0.12)               # interval 4: deeper
    oracle = make_oracle(g, adaptive=True)
    Fo, go = oracle.E_cost(x)
    neval = np.array(oracle.neval)
    assert (neval[:, 2] > 63).sum() >= 2          # the reference's quadrature did refine
    if ftile:
        monkeypatch.setenv("MFVGP_FTILE", ftile)
    fe = make_free_energy(g)
    monkeypatch.delenv("MFVGP_FTILE", raising=False)
    assert fe.describe()["path"] == "walk" and int(fe.describe()["ntiles"]) == ntiles
    F, grad = fe.E_cost(x)
    assert fe.last_status & 8                     # MFVGP_ST_REFINED
    info = fe.refine_info()
    assert np.array_equal(21 + 42 * np.maximum(info[:, 0], 1), neval[:, 0])
    assert np.array_equal(21 + 42 * np.maximum(info[:, 1], 1), neval[:, 2])
    assert abs(F - Fo) <= 1e-10 * abs(Fo), (F, Fo)
    assert rel_err(grad, go) <= 1e-10, rel_err(grad, go)
    assert worst_elem_err(grad, go) <= 1e-10, worst_elem_err(grad, go)
    # never-integrated columns (free_energy.py:711-712) are exact zeros in both
    assert np.array_equal(grad == 0.0, go == 0.0)
