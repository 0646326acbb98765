"""
CPU tests (no GPU): the numpy oracle against the golden vectors produced by the
UNMODIFIED reference (tests/golden/*.npz, generator
oracle/ref_harness/make_golden.py), and -- where /root/reference exists --
against the reference itself, live.
"""
import numpy as np
import pytest

from helpers import (EVAL_CASES, REFINE_CASES, SCG_BRANCH_CASES, SCG_CASES, load_golden,
                     make_oracle, rel_err)

RTOL = 1.0e-12


@pytest.mark.parametrize("name", [c for c in EVAL_CASES if c != "L96_500p"] + REFINE_CASES)
def test_oracle_adaptive_matches_reference(name):
    g = load_golden("eval", name)
    o = make_oracle(g, adaptive=True)
    F, grad = o.E_cost(g["x0"])
    assert abs(F - g["F"]) <= RTOL * abs(g["F"])
    assert rel_err(grad, g["grad"]) <= RTOL
    # identical quad_vec refinement decisions, interval by interval
    assert np.array_equal(np.array(o.neval), g["neval"])


@pytest.mark.parametrize("name", EVAL_CASES)
def test_oracle_fixed_rule_equals_reference_when_not_refined(name):
    """neval == 63 everywhere => quad_vec's answer is the fixed 42-node rule."""
    g = load_golden("eval", name)
    assert np.all(g["neval"] == 63)
    o = make_oracle(g, adaptive=False)
    F, grad = o.E_cost(g["x0"])
    assert abs(F - g["F"]) <= RTOL * abs(g["F"])
    assert rel_err(grad, g["grad"]) <= RTOL
    D, M = o.dim_D, o.num_M
    mp = g["x0"][:o.num_mp].reshape(D, 3 * M + 4)
    sp = np.exp(g["x0"][o.num_mp:].reshape(D, 2 * M + 3))
    Esde, dm, ds = o.E_sde(mp, sp)
    assert abs(Esde - g["Esde"]) <= RTOL * abs(g["Esde"])
    assert rel_err(dm, g["dEsde_dm"]) <= RTOL and rel_err(ds, g["dEsde_ds"]) <= RTOL
    E0, _, _ = o.E_kl0(mp[:, 0], sp[:, 0])
    Eobs, _, _ = o.E_obs(mp[o.ix_ikm], sp[o.ix_iks])
    assert abs(E0 - g["E0"]) <= RTOL * abs(g["F"])
    assert abs(Eobs - g["Eobs"]) <= RTOL * abs(g["F"])
    assert abs(o.E_cost(g["x0"], output_gradients=False) - g["F_only"]) <= RTOL * abs(g["F"])


@pytest.mark.parametrize("name", REFINE_CASES)
def test_fixed_rule_is_not_enough_when_quad_vec_refines(name):
    """Documents why the adaptive path exists."""
    g = load_golden("eval", name)
    assert g["neval"].max() > 63
    assert np.all(g["neval"][:, 1] == 63)        # the dM integrand never refines
    _, grad = make_oracle(g, adaptive=False).E_cost(g["x0"])
    assert rel_err(grad, g["grad"]) > 1.0e-10


def test_gradient_columns_never_integrated_are_obs_only():
    """free_energy.py:168,490,719: pieces M-1 and M get no SDE gradient."""
    g = load_golden("eval", "L96_40")
    D, M = 40, g["obs_times"].size
    gm = g["grad"][:D * (3 * M + 4)].reshape(D, 3 * M + 4)
    for c in (3 * M - 2, 3 * M - 1, 3 * M + 1, 3 * M + 2, 3 * M + 3):
        assert np.all(gm[:, c] == 0.0)


@pytest.mark.parametrize("name", ["OU"] + SCG_BRANCH_CASES)
def test_oracle_scg_matches_reference_trajectory(name):
    g = load_golden("scg", name)
    o = make_oracle(g, adaptive=True)
    ev = {}
    x, fx, stats = o.find_minimum(g["x0"].copy(), max_iter=int(g["max_iter"]), events=ev)
    if name in SCG_BRANCH_CASES:
        # the reference's run rejected trial points, reset a non-descent direction and
        # restarted after dim_x successes -- and the restatement takes every one of them
        assert int(g["func_eval"]) < 3 * int(g["nit"]) + 1
        assert (ev["reject"], ev["mu>=0"], ev["delta<=0"], ev["restart"]) == (
            int(g["n_reject"]), int(g["n_mu_ge0"]), int(g["n_delta_le0"]), int(g["n_restart"]))
        assert ev["reject"] >= 2
        assert rel_err(stats["dfx"][:stats["nit"]], g["dfx_hist"][:stats["nit"]]) <= 1.0e-7
    assert stats["nit"] == int(g["nit"]) and stats["func_eval"] == int(g["func_eval"])
    nit = stats["nit"]
    assert rel_err(stats["fx"][:nit], g["fx_hist"][:nit]) <= 1.0e-9
    assert rel_err(x, g["x_opt"]) <= 1.0e-7


def test_reference_goldens_cover_every_rare_scg_branch():
    """rejection (scaled_cg.py:254-279), mu >= 0 (:194-197), delta <= 0 (:233-236) and the restart
    after dim_x successes (:361-364) each occur in at least one run of the unmodified reference"""
    tot = {"n_reject": 0, "n_mu_ge0": 0, "n_delta_le0": 0, "n_restart": 0}
    for name in SCG_BRANCH_CASES:
        g = load_golden("scg", name)
        for k in tot:
            tot[k] += int(g[k])
    assert all(v >= 2 for v in tot.values()), tot


def test_scg_on_sphere_and_rosenbrock():
    """The reference's own SCG tests (src/tests/test_scaled_cg.py:38-113)."""
    from scipy.optimize import approx_fprime
    from oracle.mfvgp_oracle import scg_minimize
    rng = np.random.default_rng(795)
    sphere = lambda v: (np.sum(v ** 2), 2.0 * v)
    x, fx, _ = scg_minimize(sphere, rng.random(5), max_it=1000)
    assert np.allclose(fx, 0.0)

    def rosen(v):
        f = lambda z: (1.0 - z[0]) ** 2 + 100.0 * (z[1] - z[0] ** 2) ** 2
        return f(v), approx_fprime(v, f, 1.0e-8)
    x, fx, _ = scg_minimize(rosen, rng.random(2), max_it=1000)
    assert np.all(np.abs(x - 1.0) < 1.0e-3)


@pytest.mark.reference
def test_oracle_against_live_reference():
    from oracle.ref_harness import harness
    if not harness.available():
        pytest.skip("/root/reference not present (GPU box)")
    from oracle.ref_harness import setups
    s = setups.make_l96(D=6, tf=1.5, seed_pts=23)
    fe = setups.free_energy_of(s, n_jobs=1)
    F, grad = fe.E_cost(s["x0"])
    g = dict(system="L96", sigma=s["sde"].sigma, theta=s["sde"].theta, mu0=s["mu0"],
             tau0=s["tau0"], obs_times=s["obs_times"], obs_values=s["obs_values"],
             obs_noise=s["obs_noise"], obs_mask=s["obs_mask"])
    Fo, go = make_oracle(g).E_cost(s["x0"])
    assert abs(F - Fo) <= RTOL * abs(F) and rel_err(go, grad) <= RTOL
