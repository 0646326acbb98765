"""
Lorenz 96 sample paths.  Lorenz 96 is chaotic, so the bar is
BIT-identity with the reference's arithmetic given the same noise:

CPU: the numpy oracle against trajectories produced by the unmodified reference
(Lorenz96.make_trajectory + collect_obs, tests/golden/trajectory_*.npz).
GPU: csrc/trajectory.cuh through the C ABI against the same goldens and the oracle.
"""
from pathlib import Path

import numpy as np
import pytest

GOLDEN = Path(__file__).resolve().parent / "golden"
CASES = ["D4", "D40", "D67"]


def _load(name):
    z = np.load(GOLDEN / f"trajectory_{name}.npz")
    return {k: z[k] for k in z.files}


def _n_obs(g):
    return int(g["n_obs"]) if g["n_obs"].ndim == 0 else [int(i) for i in g["n_obs"]]


@pytest.mark.parametrize("name", CASES)
def test_oracle_is_bit_identical_to_reference(name):
    from oracle import trajectory_oracle as O
    g = _load(name)
    tk, x = O.make_trajectory(float(g["sigma"]), float(g["theta"]), int(g["D"]), float(g["t0"]),
                              float(g["tf"]), float(g["dt"]), r_seed=int(g["seed"]))
    assert np.array_equal(tk, g["tk"]) and np.array_equal(x, g["x"])


@pytest.mark.parametrize("name", CASES)
def test_collect_obs_follows_reference(name):
    import meanfieldvargp_b200 as mf
    g = _load(name)
    mask = [bool(b) for b in g["mask"]]
    obs_t, obs_y = mf.collect_obs(g["tk"], g["x"], _n_obs(g), mask)
    assert list(obs_t) == [int(i) for i in g["obs_t"]]
    assert np.array_equal(obs_y, g["obs_y"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_path_is_bit_identical_to_reference(name):
    import meanfieldvargp_b200 as mf
    g = _load(name)
    tk, x = mf.simulate_l96(float(g["sigma"]), float(g["theta"]), int(g["D"]), float(g["t0"]),
                            float(g["tf"]), float(g["dt"]), r_seed=int(g["seed"]))
    assert np.array_equal(tk.cpu().numpy(), g["tk"])
    assert np.array_equal(x.cpu().numpy(), g["x"])
    obs_t, obs_y = mf.collect_obs(tk, x, _n_obs(g), [bool(b) for b in g["mask"]])
    assert np.array_equal(obs_y.cpu().numpy(), g["obs_y"])


@pytest.mark.gpu
@pytest.mark.parametrize("D", [5, 33, 500, 1100])
def test_gpu_path_matches_oracle_given_noise(D):
    """explicit noise, vector sigma, more dimensions than threads (D = 1100 > 1024)"""
    import meanfieldvargp_b200 as mf
    from oracle import trajectory_oracle as O
    rng = np.random.default_rng(D)
    sigma = 5.0 + 10.0 * rng.random(D)
    tk = np.arange(0.0, 0.2 + 0.002, 0.002)
    noise = rng.standard_normal((D, tk.size))
    _, xo = O.make_trajectory(sigma, 8.0, D, 0.0, 0.2, 0.002, noise=noise)
    _, x = mf.simulate_l96(sigma, 8.0, D, 0.0, 0.2, 0.002, noise=noise)
    assert np.array_equal(x.cpu().numpy(), xo)


@pytest.mark.gpu
def test_trajectory_feeds_the_free_energy_path():
    """notebook flow: simulate, observe, build x0 at the support indices, evaluate E_cost"""
    import meanfieldvargp_b200 as mf
    D, dt = 40, 0.001
    tk, x = mf.simulate_l96(40.0, 8.0, D, 0.0, 1.0, dt, r_seed=3)
    mask = [bool(i % 2 == 0) for i in range(D)]
    obs_t, obs_y = mf.collect_obs(tk, x, 4, mask)
    rng = np.random.default_rng(0)
    obs_val = obs_y.cpu().numpy() + np.sqrt(2.0) * rng.standard_normal(obs_y.shape)
    mp_idk, sp_idk = mf.support_indices(obs_t, len(tk))
    mp = x.cpu().numpy().T[:, mp_idk]
    sp = 4.0 + 2.0 * rng.random((D, len(sp_idk)))
    sde = mf.Lorenz96(40.0, 8.0, dim_D=D)
    sde.time_window = tk.cpu().numpy()
    fe = mf.FreeEnergy(sde, 10.0 * np.ones(D), 4.0 * np.ones(D), tk.cpu().numpy()[obs_t], obs_val,
                       2.0 * np.ones(sum(mask)), mask)
    F, grad = fe.E_cost(mf.pack(mp, sp))
    assert np.isfinite(F) and np.all(np.isfinite(grad))


@pytest.mark.gpu
def test_blow_up_raises_like_reference():
    import meanfieldvargp_b200 as mf
    with pytest.raises(RuntimeError, match="Invalid state vector"):
        mf.simulate_l96(40.0, 8.0, 40, 0.0, 2.0, 0.05, r_seed=575)     # dt far too large


# ---- DoubleWell / OrnsteinUhlenbeck / Lorenz63 (double_well.py:40-91, ornstein_uhlenbeck.py:38-79,
# ---- lorenz_63.py:102-181): same bar, bit identity with the unmodified reference -----------------
SMALL_CASES = ["DW", "DW_coarse", "OU", "L63", "L63_scalar_sigma"]


def _small_oracle(g):
    from oracle import trajectory_oracle as O
    fn = {"DoubleWell": O.make_trajectory_dw, "OrnsteinUhlenbeck": O.make_trajectory_ou,
          "Lorenz63": O.make_trajectory_l63}[str(g["cls"])]
    sigma = g["sigma"] if g["sigma"].ndim else float(g["sigma"])
    theta = g["theta"] if g["theta"].ndim else float(g["theta"])
    return fn(sigma, theta, float(g["t0"]), float(g["tf"]), float(g["dt"]), r_seed=int(g["seed"]))


@pytest.mark.parametrize("name", SMALL_CASES)
def test_small_system_oracle_is_bit_identical_to_reference(name):
    g = _load(name)
    tk, x = _small_oracle(g)
    assert np.array_equal(tk, g["tk"]) and np.array_equal(x, g["x"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", SMALL_CASES)
def test_small_system_gpu_path_is_bit_identical_to_reference(name):
    """through the reference's own call sequence: Class(sigma, theta, r_seed).make_trajectory,
    .collect_obs"""
    import meanfieldvargp_b200 as mf
    g = _load(name)
    sigma = g["sigma"] if g["sigma"].ndim else float(g["sigma"])
    theta = g["theta"] if g["theta"].ndim else float(g["theta"])
    sde = getattr(mf, str(g["cls"]))(sigma, theta, r_seed=int(g["seed"]))
    with pytest.raises(NotImplementedError, match="Sample path"):
        sde.collect_obs(2)
    sde.make_trajectory(float(g["t0"]), float(g["tf"]), float(g["dt"]))
    assert np.array_equal(sde.time_window, g["tk"])
    assert sde.sample_path.shape == g["x"].shape and np.array_equal(sde.sample_path, g["x"])
    obs_t, obs_y = sde.collect_obs(int(g["n_obs"]))
    assert list(obs_t) == [int(i) for i in g["obs_t"]] and np.array_equal(obs_y, g["obs_y"])


@pytest.mark.gpu
def test_l96_class_make_trajectory_matches_reference():
    import meanfieldvargp_b200 as mf
    g = _load("D40")
    sde = mf.Lorenz96(float(g["sigma"]), float(g["theta"]), dim_D=int(g["D"]), r_seed=int(g["seed"]))
    sde.make_trajectory(float(g["t0"]), float(g["tf"]), float(g["dt"]))
    assert np.array_equal(sde.sample_path, g["x"])
    obs_t, obs_y = sde.collect_obs(_n_obs(g), [bool(b) for b in g["mask"]])
    assert np.array_equal(obs_y, g["obs_y"])


@pytest.mark.gpu
@pytest.mark.parametrize("system", ["DW", "OU", "L63"])
def test_small_system_ensemble_matches_oracle_path_by_path(system):
    """npaths independent paths in one launch (one thread each) == the oracle run per child seed"""
    import meanfieldvargp_b200 as mf
    from oracle import trajectory_oracle as O
    npaths, seed = 300, 17                       # more than two CTAs of 128
    sim, orc, sigma, theta = {
        "DW": (mf.simulate_dw, O.make_trajectory_dw, 0.8, 1.0),
        "OU": (mf.simulate_ou, O.make_trajectory_ou, 0.5, [2.0, 1.5]),
        "L63": (mf.simulate_l63, O.make_trajectory_l63, [40.0, 30.0, 20.0], [10.0, 28.0, 2.66667]),
    }[system]
    tk, x = sim(sigma, theta, 0.0, 0.5, 0.005, r_seed=seed, npaths=npaths)
    x = x.cpu().numpy()
    assert x.shape[:2] == (npaths, tk.numel())
    children = np.random.SeedSequence(seed).spawn(npaths)
    for p in (0, 1, 127, 128, 255, 299):
        _, xo = orc(sigma, theta, 0.0, 0.5, 0.005, rng=np.random.default_rng(children[p]))
        assert np.array_equal(x[p], xo), (system, p)


@pytest.mark.gpu
def test_l63_blow_up_raises_like_reference():
    import meanfieldvargp_b200 as mf
    with pytest.raises(RuntimeError, match="Invalid state vector"):
        mf.simulate_l63(40.0, [10.0, 28.0, 2.66667], 0.0, 40.0, 0.2, r_seed=1)      # dt far too large


@pytest.mark.gpu
def test_dw_notebook_flow_on_simulated_data():
    """example_DW.ipynb: simulate, observe twice per time unit, optimise from the path itself"""
    import meanfieldvargp_b200 as mf
    sde = mf.DoubleWell(0.8, 1.0, r_seed=575)
    sde.make_trajectory(0.0, 10.0, 0.001)
    obs_t, obs_y = sde.collect_obs(2)
    rng = np.random.default_rng(1)
    obs_val = obs_y + np.sqrt(0.04) * rng.standard_normal(obs_y.shape)
    mp_idk, sp_idk = mf.support_indices(obs_t, len(sde.time_window))
    mp = sde.sample_path[mp_idk][None, :]
    sp = np.log(0.2 + 0.1 * rng.random((1, len(sp_idk))))
    fe = mf.FreeEnergy(sde, 1.0, 0.2, sde.time_window[obs_t], obs_val, 0.04, [True])
    x0 = np.concatenate([mp.ravel(), sp.ravel()])
    F0, _ = fe.E_cost(x0)
    _, Fx, stats = fe.find_minimum(x0, max_iter=30)
    assert np.isfinite(F0) and Fx < F0
