"""
Gradients with respect to the drift parameters theta and the diffusion coefficients Sigma
(hooks in the reference: free_energy.py:241-279, :745-748).

Goldens (tests/golden/params_*.npz, generator oracle/ref_harness/make_golden_params.py) hold
(1) SymPy derivatives of the reference's own decoded energy expressions, integrated with
quad_vec as the reference integrates, and (2) central differences of the UNMODIFIED reference's
E_cost through its setters.  CPU tests pin the numpy oracle to them; GPU tests compare
mfvgp_eparam with the goldens and with the oracle at sizes the goldens do not reach.
"""
import numpy as np
import pytest

from helpers import load_golden, make_free_energy, make_oracle, rel_err, synthetic_l96

PARAM_CASES = ["DW", "OU", "L63_T5", "L96_8", "L96_40", "L96_67p"]
REFINING = ["DW_refine", "L96_6_refine"]      # quad_vec bisects further on some intervals
RTOL = 1.0e-10          # north_star tolerance
FD_RTOL = 1.0e-7        # central differences of the reference: step 1e-5, rounding ~1e-9


def _oracle_params(g, x, adaptive=True):
    from oracle.param_oracle import param_gradients
    return param_gradients(make_oracle(g, adaptive=adaptive), x, adaptive=adaptive)


@pytest.mark.parametrize("name", PARAM_CASES + REFINING)
def test_oracle_matches_sympy_of_reference_expressions(name):
    g = load_golden("params", name)
    dth, dsg = _oracle_params(g, g["x0"])
    assert rel_err(dth, g["dF_dtheta"]) <= 1e-12
    assert rel_err(dsg, g["dF_dsigma"]) <= 1e-12


@pytest.mark.parametrize("name", PARAM_CASES)
def test_golden_sympy_agrees_with_finite_differences_of_the_reference(name):
    g = load_golden("params", name)
    assert rel_err(g["dF_dtheta"], g["fd_dtheta"]) <= FD_RTOL
    idx = g["fd_sigma_idx"]
    assert rel_err(g["dF_dsigma"][idx], g["fd_dsigma"]) <= FD_RTOL


@pytest.mark.parametrize("name", PARAM_CASES)
def test_fixed_rule_is_quad_vecs_answer_on_the_goldens(name):
    g = load_golden("params", name)
    assert np.all(g["neval"] == 63)
    dth, dsg = _oracle_params(g, g["x0"], adaptive=False)
    assert rel_err(dth, g["dF_dtheta"]) <= 1e-12 and rel_err(dsg, g["dF_dsigma"]) <= 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("name", PARAM_CASES)
def test_gpu_param_gradients_match_reference_goldens(name):
    g = load_golden("params", name)
    fe = make_free_energy(g)
    dth, dsg = fe.grad_params(g["x0"])
    assert fe.last_status == 0                   # fixed rule converged everywhere
    assert dth.shape == g["dF_dtheta"].shape and dsg.shape == g["dF_dsigma"].shape
    assert rel_err(dth, g["dF_dtheta"]) <= RTOL, rel_err(dth, g["dF_dtheta"])
    assert rel_err(dsg, g["dF_dsigma"]) <= RTOL, rel_err(dsg, g["dF_dsigma"])
    assert rel_err(dth, g["fd_dtheta"]) <= FD_RTOL
    assert rel_err(dsg[g["fd_sigma_idx"]], g["fd_dsigma"]) <= FD_RTOL


@pytest.mark.gpu
@pytest.mark.parametrize("name", REFINING)
def test_gpu_param_gradients_where_quad_vec_refines(name):
    """Variance support points that differ by 10-300x: quad_vec (limit=100, epsrel=1e-6) bisects
    these integrands up to 7 times.  The reference has no parameter-gradient integrand whose
    bisection order could be mirrored, so the kernel switches such cells to a composite 16-panel
    GK21 rule (exact to rounding) and raises MFVGP_ST_REFINED: agreement is then bounded by
    quad_vec's own accuracy (epsrel 1e-6; observed ~1e-9), not by 1e-10."""
    g = load_golden("params", name)
    assert g["neval"].max() > 63
    fe = make_free_energy(g)
    dth, dsg = fe.grad_params(g["x0"])
    assert fe.last_status & 8
    assert rel_err(dth, g["dF_dtheta"]) <= 1e-7, rel_err(dth, g["dF_dtheta"])
    assert rel_err(dsg, g["dF_dsigma"]) <= 1e-7, rel_err(dsg, g["dF_dsigma"])
    assert rel_err(dth, g["fd_dtheta"]) <= 1e-6
    assert rel_err(dsg[g["fd_sigma_idx"]], g["fd_dsigma"]) <= 1e-6
    # the fixed rule alone is NOT enough here
    from oracle.param_oracle import param_gradients
    _, dsg_fixed = param_gradients(make_oracle(g, adaptive=False), g["x0"], adaptive=False)
    assert rel_err(dsg_fixed, g["dF_dsigma"]) > 10 * rel_err(dsg, g["dF_dsigma"])


@pytest.mark.gpu
def test_gpu_param_gradients_tensor_path_and_setters():
    """tensor in -> tensors out; the values follow the sde_theta / sde_sigma setters
    (free_energy.py:241-279), checked by central differences of our own E_cost"""
    import torch
    g = load_golden("params", "L96_8")
    fe = make_free_energy(g)
    x = torch.tensor(g["x0"], dtype=torch.float64, device="cuda")
    dth, dsg = fe.grad_params(x)
    assert dth.is_cuda and rel_err(dth.cpu().numpy(), g["dF_dtheta"]) <= RTOL
    assert rel_err(dsg.cpu().numpy(), g["dF_dsigma"]) <= RTOL
    th0, sg0 = fe.sde_theta.copy(), fe.sde_sigma.copy()
    fe.sde_theta = th0 * 1.1
    sg1 = sg0.copy()
    sg1[3] *= 0.8
    fe.sde_sigma = sg1
    dth1, dsg1 = fe.grad_params(g["x0"])
    h = 1e-5
    F = []
    for s in (+1, -1):
        fe.sde_theta = th0 * 1.1 + s * h
        F.append(fe.E_cost(g["x0"], output_gradients=False))
    assert abs((F[0] - F[1]) / (2 * h) - dth1[0]) <= 1e-7 * abs(dth1[0])
    fe.sde_theta = th0 * 1.1
    F = []
    for s in (+1, -1):
        sg = sg1.copy()
        sg[3] += s * h
        fe.sde_sigma = sg
        F.append(fe.E_cost(g["x0"], output_gradients=False))
    assert abs((F[0] - F[1]) / (2 * h) - dsg1[3]) <= 1e-7 * abs(dsg1[3])


@pytest.mark.gpu
@pytest.mark.parametrize("D_,M_", [(500, 16), (1000, 7), (8192, 5)])
def test_gpu_param_gradients_match_oracle_at_size(D_, M_):
    """the north-star size and the 68-tile size against the neighbour-sparse oracle"""
    g = synthetic_l96(D_, M_, seed=31 + D_)
    g["sigma"] = g["sigma"] * (0.5 + np.random.default_rng(D_).random(D_))     # non-uniform Sigma
    fe = make_free_energy(g)
    dth, dsg = fe.grad_params(g["x0"])
    oth, osg = _oracle_params(g, g["x0"], adaptive=False)
    assert rel_err(dth, oth) <= RTOL, rel_err(dth, oth)
    assert rel_err(dsg, osg) <= RTOL, rel_err(dsg, osg)
