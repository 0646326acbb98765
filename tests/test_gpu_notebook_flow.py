"""
The whole of examples/example_L96.ipynb on this package alone: sample path and observations on the
device, the notebook's initial path and state, FreeEnergy.find_minimum with check_gradients=True
(batched on the device), posterior reconstruction.  A drop-in check of the steps around the path
working TOGETHER; the numerical bars of the individual steps live in their own test files.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_example_l96_notebook_end_to_end(capsys):
    import meanfieldvargp_b200 as mf
    D, t0, tf, dt = 40, 0.0, 5.0, 0.001
    sde = mf.Lorenz96(40.0, 8.0, dim_D=D, r_seed=575)            # example_L96.ipynb:90-133
    sde.make_trajectory(t0, tf, dt)
    obs_t, obs_y = sde.collect_obs(4)
    rng = np.random.default_rng(1)
    obs_noise = 2.0 * np.ones(D)
    obs_val = obs_y + np.sqrt(obs_noise) * sde.rng.standard_normal(obs_y.shape)   # as the notebook
    M = len(obs_t)
    assert M == 20 and obs_y.shape == (M, D)

    path0 = mf.initial_path(sde.time_window, obs_t, obs_val)
    x0, mp_idk, sp_idk = mf.initial_state(path0, obs_t, rng=rng)
    assert x0.size == D * (5 * M + 7)

    fe = mf.FreeEnergy(sde, 10.0 * np.ones(D), 4.0 * np.ones(D), sde.time_window[obs_t], obs_val,
                       obs_noise, n_jobs=8)                       # :363-367, n_jobs accepted
    F0, _ = fe.E_cost(x0)
    x, Fx, stats = fe.find_minimum(x0, max_iter=60, check_gradients=True, verbose=True)
    out = capsys.readouterr().out
    assert out.count("Grad-Check |") == 2 and "SCG: Optimization started" in out
    errs = [float(line.split("=")[1]) for line in out.splitlines() if line.startswith(" > Error")]
    assert len(errs) == 2 and max(errs) < 1.0e-2                  # forward-difference noise level
    assert Fx < 0.7 * F0 and stats["nit"] == 60

    mean_pts, vars_pts = mf.unpack(x, D, M)
    tt, mt, st = mf.reconstruct(mean_pts, vars_pts, t0, sde.time_window[obs_t], tf, num=25)
    assert mt.shape == (D, (M + 1) * 25 + 1) and np.all(st > 0.0)
    truth = np.array([np.interp(tt, sde.time_window, sde.sample_path[:, i]) for i in range(D)])
    rmse_post = float(np.sqrt(np.mean((mt - truth) ** 2)))
    prior0 = np.array([np.interp(tt, sde.time_window, path0[i]) for i in range(D)])
    rmse_init = float(np.sqrt(np.mean((prior0 - truth) ** 2)))
    assert rmse_post < rmse_init                                  # smoothing improved on the splines
