"""
Device SCG (mfvgp_scg) against the reference's SCG.minimize trajectories
(golden vectors from the unmodified reference, scaled_cg.py:107-386).

SCG is a branchy iteration: the bar is the identical branch sequence (same
nit and func_eval, same accept/reject pattern visible in fx) with the energy
history matching to 1e-8 relative and the final iterate to 1e-6.
"""
import numpy as np
import pytest

from helpers import SCG_CASES, load_golden, make_free_energy, make_oracle, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", SCG_CASES)
def test_scg_matches_reference_trajectory(name, capsys):
    g = load_golden("scg", name)
    fe = make_free_energy(g)
    x, fx, stats = fe.find_minimum(g["x0"].copy(), max_iter=int(g["max_iter"]))
    out = capsys.readouterr().out
    assert "Elapsed time [" in out and "Done!" in out
    assert stats["nit"] == int(g["nit"])
    assert stats["func_eval"] == int(g["func_eval"])
    nit = stats["nit"]
    assert rel_err(stats["fx"][:nit], g["fx_hist"][:nit]) <= 1.0e-8
    assert rel_err(stats["dfx"][:nit], g["dfx_hist"][:nit]) <= 1.0e-5
    assert abs(fx - float(g["fx_opt"])) <= 1.0e-8 * abs(float(g["fx_opt"]))
    assert rel_err(x, g["x_opt"]) <= 1.0e-6


def test_scg_tensor_in_place_and_floor():
    import torch
    g = load_golden("scg", "L96_8")
    fe = make_free_energy(g)
    x0 = torch.tensor(g["x0"], dtype=torch.float64, device="cuda")
    # max_iter is floored at 10 (free_energy.py:815)
    x, fx, stats = fe.find_minimum(x0, max_iter=3)
    assert stats["fx"].size == 10 and x.is_cuda
    assert fx <= float(fe.E_cost(x0)[0].item())
    assert torch.equal(x0, torch.tensor(g["x0"], dtype=torch.float64, device="cuda"))


def test_scg_decreases_energy_full_size():
    """North-star size: D=500 synthetic; SCG must reduce F monotonically in fx."""
    from helpers import synthetic_l96
    g = synthetic_l96(500, 16, seed=11)
    fe = make_free_energy(g)
    F0 = fe.E_cost(g["x0"], output_gradients=False)
    x, fx, stats = fe.find_minimum(g["x0"], max_iter=12)
    assert fx < F0
    hist = stats["fx"][:stats["nit"]]
    assert np.all(np.diff(hist) <= 1e-9 * abs(F0))
