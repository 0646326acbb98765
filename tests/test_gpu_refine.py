"""
Adaptive quadrature on the device (refine.cuh) against the reference on inputs
where scipy's quad_vec refines beyond its first bisection:
same refinement decisions (evaluation counts) and results to 1e-10.
"""
import numpy as np
import pytest

from helpers import REFINE_CASES, load_golden, make_free_energy, make_oracle, rel_err

pytestmark = pytest.mark.gpu
RTOL = 1.0e-10


def _neval_from_info(info):
    """bisections -> scipy neval; unflagged intervals took the fixed rule (63)."""
    return np.where(info > 0, 21 + 42 * info, 63)


@pytest.mark.parametrize("name", REFINE_CASES)
def test_refined_matches_reference(name):
    g = load_golden("eval", name)
    fe = make_free_energy(g)
    F, grad = fe.E_cost(g["x0"])
    info = fe.refine_info()
    assert fe.last_status & 8, "refinement should have been flagged"
    assert np.array_equal(_neval_from_info(info[:, 0]), g["neval"][:, 0])
    assert np.array_equal(_neval_from_info(info[:, 1]), g["neval"][:, 2])
    assert abs(F - g["F"]) <= RTOL * abs(g["F"])
    assert rel_err(grad, g["grad"]) <= RTOL
    D, M = fe.dim_D, fe.num_M
    mp = g["x0"][:fe.num_mp].reshape(D, 3 * M + 4)
    sp = np.exp(g["x0"][fe.num_mp:].reshape(D, 2 * M + 3))
    Esde, dm, ds = fe.E_sde(mp, sp)
    assert abs(Esde - g["Esde"]) <= RTOL * abs(g["Esde"])
    assert rel_err(dm, g["dEsde_dm"]) <= RTOL
    assert rel_err(ds, g["dEsde_ds"]) <= RTOL
    # value-only call: only the energy integrand is integrated (free_energy.py:409)
    assert abs(fe.E_cost(g["x0"], output_gradients=False) - g["F_only"]) <= RTOL * abs(g["F"])


def test_unrefined_inputs_take_fixed_rule():
    g = load_golden("eval", "L96_40")
    fe = make_free_energy(g)
    fe.E_cost(g["x0"])
    assert not (fe.last_status & 8)
    assert np.all(fe.refine_info() == 0)


@pytest.mark.parametrize("system,seed,lo", [("L63", 3, -1.0), ("OU", 4, -1.0), ("DW", 5, -1.5)])
def test_random_rough_variances_against_oracle(system, seed, lo):
    """Random inputs with strongly varying variances, checked against scipy live."""
    g = dict(load_golden("eval", system))
    rng = np.random.default_rng(seed)
    D, M = g["sigma"].size, g["obs_times"].size
    n_mp = D * (3 * M + 4)
    x = g["x0"].copy()
    x[n_mp:] = np.log(10.0 ** rng.uniform(lo, 0.0, size=D * (2 * M + 3)))
    fe = make_free_energy(g)
    F, grad = fe.E_cost(x)
    o = make_oracle(g, adaptive=True)
    Fo, go = o.E_cost(x)
    nev = np.array(o.neval)
    assert 63 < nev.max() < 2000          # refines, but never hits limit=100
    info = fe.refine_info()
    assert np.array_equal(_neval_from_info(info[:, 0]), nev[:, 0])
    assert np.array_equal(_neval_from_info(info[:, 1]), nev[:, 2])
    assert abs(F - Fo) <= RTOL * abs(Fo)
    assert rel_err(grad, go) <= RTOL


def test_limit_is_reported_not_hidden():
    """
    quad_vec(limit=100) gives up on pathological inputs; the device path stops at
    the same panel budget and says so in the status word.  (Heap order then
    depends on 1-ulp ties, so only the flag and finiteness are asserted.)
    """
    g = dict(load_golden("eval", "L63"))
    rng = np.random.default_rng(3)
    D, M = 3, g["obs_times"].size
    x = g["x0"].copy()
    x[D * (3 * M + 4):] = np.log(10.0 ** rng.uniform(-2.0, 0.0, size=D * (2 * M + 3)))
    fe = make_free_energy(g)
    F, grad = fe.E_cost(x)
    assert fe.last_status & 16 and fe.last_status & 8
    assert np.isfinite(F) and np.all(np.isfinite(grad))
    assert fe.refine_info().max() >= 99


@pytest.mark.parametrize("D,group", [(130, 3), (1100, None), (2100, None)])
def test_refine_groups_match_oracle(D, group, monkeypatch):
    """
    At large D a flagged interval is shared by a group of CTAs (refine.cuh GroupCtx: rows dealt
    over the group, reductions through global memory + group barrier).  Same decisions and
    values as scipy's quad_vec; `group` forces a group size on a small problem.
    """
    from helpers import synthetic_l96
    M = 4
    g = synthetic_l96(D, M, seed=D)
    x = g["x0"].copy()
    sp = x[D * (3 * M + 4):].reshape(D, 2 * M + 3)
    sp[:, [1, 4]] = np.log(0.2)                 # steep variance triples in two intervals
    if group:
        monkeypatch.setenv("MFVGP_REFINE_GROUP", str(group))
    monkeypatch.setenv("MFVGP_PATH", "walk")     # refine_kernel (the coop path refines in-kernel)
    fe = make_free_energy(g)
    monkeypatch.delenv("MFVGP_PATH")
    monkeypatch.delenv("MFVGP_REFINE_GROUP", raising=False)
    F, grad = fe.E_cost(x)
    info = fe.refine_info()
    assert info.max() >= 2
    o = make_oracle(g, adaptive=True)
    Fo, go = o.E_cost(x)
    nev = np.array(o.neval)
    assert np.array_equal(_neval_from_info(info[:, 0]), nev[:, 0])
    assert np.array_equal(_neval_from_info(info[:, 1]), nev[:, 2])
    assert abs(F - Fo) <= RTOL * abs(Fo) and rel_err(grad, go) <= RTOL
    assert not (fe.last_status & 32)


def test_refine_groups_of_16_equal_single_cta(monkeypatch):
    """D = 8192: groups of 16 CTAs against one CTA per item (same kernel, MFVGP_REFINE_GROUP=1)."""
    import torch
    from helpers import synthetic_l96_device
    D, M = 8192, 41
    g = synthetic_l96_device(torch, D, M, seed=5)
    x = g["x0"].clone()
    sp = x[D * (3 * M + 4):].view(D, 2 * M + 3)
    sp[:, [5, 21, 40]] = float(np.log(0.25))
    fe16 = make_free_energy(g)
    F16, g16 = fe16.E_cost(x)
    info16 = fe16.refine_info()
    monkeypatch.setenv("MFVGP_REFINE_GROUP", "1")
    fe1 = make_free_energy(g)
    monkeypatch.delenv("MFVGP_REFINE_GROUP")
    F1, g1 = fe1.E_cost(x)
    assert info16.max() >= 2 and np.array_equal(fe1.refine_info(), info16)
    assert abs(F16.item() - F1.item()) <= 1e-13 * abs(F1.item())
    assert rel_err(g16.cpu().numpy(), g1.cpu().numpy()) <= 1e-12
