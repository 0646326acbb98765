"""
Batched evaluation (mfvgp_ecost_batch / FreeEnergy.E_cost_batch): B states of one problem in two
launches.  Bar: every row is bit-identical to the single-state E_cost of that row -- including
rows the guard flags, which go through the adaptive stage -- and the batched gradient check
returns what scipy.optimize.check_grad returns (free_energy.py:805).
"""
import numpy as np
import pytest

from helpers import load_golden, make_free_energy, rel_err, synthetic_l96

pytestmark = pytest.mark.gpu


def _states(g, B, seed, scale=0.05):
    rng = np.random.default_rng(seed)
    x0 = np.asarray(g["x0"], dtype=float)
    X = x0[None, :] + scale * rng.standard_normal((B, x0.size))
    X[0] = x0
    return X


@pytest.mark.parametrize("kernel", ["coop", "split", "lean"])
@pytest.mark.parametrize("case,B", [("L96_40", 37), ("L96_500p", 9), ("L96_5_irregular", 130),
                                    ("L96_67p", 70)])
def test_batch_rows_equal_single_evaluations(case, B, kernel, monkeypatch):
    """main kernel of the batch = the cooperative kernel of the single evaluation (bit-identical
    rows) or the one-thread-per-cell split kernel (same sums in another order: rounding)"""
    import torch
    monkeypatch.setenv("MFVGP_BATCH_KERNEL", kernel)
    g = load_golden("eval", case)
    fe = make_free_energy(g)
    assert fe.supports_batch()
    X = torch.as_tensor(_states(g, B, seed=B), device="cuda")
    F, G = fe.E_cost_batch(X, chunk=32)                    # several chunks, a ragged last one
    Fo = fe.E_cost_batch(X, output_gradients=False, chunk=32)
    assert torch.equal(F, Fo)
    for b in range(B):
        f1, g1 = fe.E_cost(X[b])
        if kernel == "coop":
            assert float(f1) == float(F[b]) and torch.equal(g1, G[b]), b
        else:
            assert abs(float(f1) - float(F[b])) <= 1.0e-12 * abs(float(f1)), b
            assert rel_err(G[b].cpu().numpy(), g1.cpu().numpy()) <= 1.0e-11, b
    # row 0 is the golden state: the reference's own numbers
    assert abs(float(F[0]) - float(g["F"])) <= 1.0e-10 * abs(float(g["F"]))
    assert rel_err(G[0].cpu().numpy(), g["grad"]) <= 1.0e-10


def test_batch_flagged_rows_take_the_adaptive_stage():
    """a state on which quad_vec bisects beyond the fixed rule, between ordinary ones"""
    import torch
    g = load_golden("eval", "L96_6_refine")
    fe = make_free_energy(g)
    X = _states(g, 5, seed=3, scale=1.0e-3)
    X[3] = g["x0"]
    F, G = fe.E_cost_batch(X)                               # numpy in -> numpy out
    assert isinstance(F, np.ndarray) and G.shape == X.shape
    assert abs(F[3] - float(g["F"])) <= 1.0e-10 * abs(float(g["F"]))
    assert rel_err(G[3], g["grad"]) <= 1.0e-10
    for b in range(5):
        f1, g1 = fe.E_cost(X[b])
        assert f1 == F[b] and np.array_equal(g1, G[b]), b


@pytest.mark.parametrize("kernel", ["coop", "split", "lean"])
def test_batched_gradient_check_equals_scipy_check_grad(kernel, capsys, monkeypatch):
    from scipy.optimize import check_grad
    monkeypatch.setenv("MFVGP_BATCH_KERNEL", kernel)
    g = load_golden("scg", "L96_8")
    fe = make_free_energy(g)
    x = np.asarray(g["x0"], dtype=float)
    ref = check_grad(lambda v: fe.E_cost(v, output_gradients=False), lambda v: fe.E_cost(v)[1], x)
    got = fe.check_grad_batched(x, chunk=100)
    # the number is the rounding noise of forward differences (1e-5 of a gradient of 1e2): the
    # cooperative kernel reproduces scipy's digits, the split kernel rounds differently
    assert abs(got - ref) <= (1.0e-6 if kernel == "coop" else 0.5) * ref
    assert got < 1.0e-3
    # and through the reference's call: find_minimum(check_gradients=True) prints it twice
    fe.find_minimum(x, max_iter=10, check_gradients=True)
    out = capsys.readouterr().out
    assert out.count("Grad-Check |") == 2 and out.count(" > Error = ") == 2


@pytest.mark.parametrize("case", ["DW", "OU", "L63", "DW_refine"])
def test_batch_of_the_small_systems(case):
    """DW / OU / L63: the batch runs the single evaluation's own kernels (guard inside), so rows are
    bit-identical; DW flags intervals often, those rows go through the adaptive stage"""
    g = load_golden("eval", case)
    fe = make_free_energy(g)
    assert fe.supports_batch()
    X = _states(g, 40, seed=7, scale=0.02)
    F, G = fe.E_cost_batch(X, chunk=16)
    for b in range(X.shape[0]):
        f1, g1 = fe.E_cost(X[b])
        assert f1 == F[b] and np.array_equal(g1, G[b]), b
    assert abs(F[0] - float(g["F"])) <= 1.0e-10 * abs(float(g["F"]))
    assert rel_err(G[0], g["grad"]) <= 1.0e-10


def test_small_system_gradient_check_goes_through_the_batch(capsys):
    """example_DW / example_OU / example_L63 call find_minimum(check_gradients=True)"""
    from scipy.optimize import check_grad
    g = load_golden("scg", "DW")
    fe = make_free_energy(g)
    x = np.asarray(g["x0"], dtype=float)
    ref = check_grad(lambda v: fe.E_cost(v, output_gradients=False), lambda v: fe.E_cost(v)[1], x)
    got = fe.check_grad_batched(x, chunk=50)
    assert abs(got - ref) <= 1.0e-6 * ref


def test_batch_is_refused_where_it_does_not_apply():
    big = synthetic_l96(256, 200, seed=1)                   # 51 000 cells: walk path
    fb = make_free_energy(big)
    assert fb.describe()["path"] == "walk" and not fb.supports_batch()
    with pytest.raises(NotImplementedError):
        fb.E_cost_batch(np.asarray(big["x0"], dtype=float)[None, :])
