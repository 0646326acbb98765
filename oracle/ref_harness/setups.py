"""
TEST INFRASTRUCTURE (oracle side): builds the benchmark/parity problems of
SURVEY.md section 8(d) with the REFERENCE's own system classes
(``make_trajectory`` / ``collect_obs``), so it only runs where
/root/reference exists.  All randomness is seeded.

Notebook sources of the parameters: ``examples/example_DW.ipynb`` (cells
8, 10, 19, 21), ``example_OU.ipynb``, ``example_L63.ipynb``,
``example_L96.ipynb``, ``example_500D.ipynb``.
"""
import numpy as np
from scipy.interpolate import splev, splrep

from . import harness


def _support_indices(obs_idk, n_t):
    """Support-point index selection, ``example_500D.ipynb`` cell 17."""
    time_space = [0, *obs_idk, n_t - 1]
    mp_idk, sp_idk = [], []
    for n in range(len(time_space) - 1):
        mp_idk.extend(np.linspace(time_space[n], time_space[n + 1], num=4,
                                  endpoint=True, dtype=int))
        sp_idk.extend(np.linspace(time_space[n], time_space[n + 1], num=3,
                                  endpoint=True, dtype=int))
    return np.unique(mp_idk), np.unique(sp_idk)


def _spline_path(sde, t0, tf, dt, obs_idk, obs_val):
    """Cubic B-spline through the noisy observations (notebook cell 14)."""
    spl_t = np.array([t0, *(np.asarray(obs_idk) * dt), tf])
    vals = np.atleast_2d(obs_val.T).T if obs_val.ndim == 1 else obs_val
    vals = vals.reshape(len(obs_idk), -1)
    out = []
    for i in range(vals.shape[1]):
        spl_v = np.array([vals[0, i], *vals[:, i], vals[-1, i]])
        out.append(splev(sde.tk, splrep(spl_t, spl_v, k=3)))
    return np.array(out)


def _finish(name, sde, t0, tf, dt, obs_idk, obs_val, obs_noise, obs_mask,
            mu0, tau0, mp, sp_log, n_jobs):
    return dict(name=name, sde=sde, dt=dt, t0=t0, tf=tf,
                obs_times=np.asarray(obs_idk) * dt, obs_values=obs_val,
                obs_noise=obs_noise, obs_mask=obs_mask, mu0=mu0, tau0=tau0,
                x0=np.concatenate((np.asarray(mp).flatten(order="C"),
                                   np.asarray(sp_log).flatten(order="C"))),
                n_jobs=n_jobs)


def make_dw(seed_pts=1, tf=10.0):
    C = harness.reference_classes()
    sde = C["DoubleWell"](0.8, 1.0, r_seed=575)
    t0, dt = 0.0, 0.001
    sde.make_trajectory(t0, tf, dt)
    obs_idk, obs_val = sde.collect_obs(2)
    obs_noise = 0.04
    obs_val = obs_val + np.sqrt(obs_noise) * sde.rng.standard_normal(obs_val.shape)
    x0 = _spline_path(sde, t0, tf, dt, obs_idk, obs_val)[0]
    mp_idk, sp_idk = _support_indices(obs_idk, len(x0))
    rng = np.random.default_rng(seed_pts)
    sp = np.log(0.2 + 0.1 * rng.random(len(sp_idk)))
    return _finish("DW", sde, t0, tf, dt, obs_idk, obs_val, obs_noise, None,
                   1.0, 0.2, x0[mp_idk], sp, 1)


def make_ou(seed_pts=1, tf=10.0):
    C = harness.reference_classes()
    sde = C["OrnsteinUhlenbeck"](0.5, [2.0, 1.0], r_seed=575)
    t0, dt = 0.0, 0.001
    sde.make_trajectory(t0, tf, dt)
    obs_idk, obs_val = sde.collect_obs(2)
    obs_noise = 0.04
    obs_val = obs_val + np.sqrt(obs_noise) * sde.rng.standard_normal(obs_val.shape)
    x0 = _spline_path(sde, t0, tf, dt, obs_idk, obs_val)[0]
    mp_idk, sp_idk = _support_indices(obs_idk, len(x0))
    rng = np.random.default_rng(seed_pts)
    sp = np.log(0.2 + 0.1 * rng.random(len(sp_idk)))
    return _finish("OU", sde, t0, tf, dt, obs_idk, obs_val, obs_noise, None,
                   1.0, 0.2, x0[mp_idk], sp, 1)


def make_l63(seed_pts=1, tf=20.0):
    C = harness.reference_classes()
    sde = C["Lorenz63"](np.array([40.0, 40.0, 40.0]),
                        np.array([10.0, 28.0, 2.66667]), r_seed=575)
    t0, dt = 0.0, 0.001
    sde.make_trajectory(t0, tf, dt)
    obs_idk, obs_val = sde.collect_obs(3)
    obs_noise = np.array([2.0, 2.0, 2.0])
    obs_val = obs_val + np.sqrt(obs_noise) * sde.rng.standard_normal(obs_val.shape)
    x0 = _spline_path(sde, t0, tf, dt, obs_idk, obs_val)
    mp_idk, sp_idk = _support_indices(obs_idk, x0.shape[1])
    rng = np.random.default_rng(seed_pts)
    sp = np.log(4.0 + 2.0 * rng.random((3, len(sp_idk))))
    return _finish("L63", sde, t0, tf, dt, obs_idk, obs_val, obs_noise, None,
                   10.0 * np.ones(3), 5.0 * np.ones(3), x0[:, mp_idk], sp, 1)


def make_l96(D=40, tf=5.0, partial=False, seed_pts=1, n_jobs=1, sigma=40.0):
    C = harness.reference_classes()
    sde = C["Lorenz96"](sigma * np.ones(D), 8.0, dim_D=D, r_seed=575)
    t0, dt = 0.0, 0.001
    sde.make_trajectory(t0, tf, dt)
    h_mask = [bool(i % 4 == 0) for i in range(D)] if partial else None
    obs_idk, obs_val = sde.collect_obs(4, h_mask)
    num_M, dim_o = obs_val.shape
    obs_noise = 2.0 * np.ones(dim_o)
    obs_val = obs_val + np.sqrt(obs_noise) * sde.rng.standard_normal((num_M, dim_o))
    mp_idk, sp_idk = _support_indices(obs_idk, len(sde.tk))
    rng = np.random.default_rng(seed_pts)
    mp = sde.xt[mp_idk].T + 0.5 * rng.standard_normal((D, len(mp_idk)))
    sp = np.log(4.0 + 2.0 * rng.random((D, len(sp_idk))))
    name = f"L96_{D}" + ("p" if partial else "")
    return _finish(name, sde, t0, tf, dt, obs_idk, obs_val, obs_noise, h_mask,
                   10.0 * np.ones(D), 4.0 * np.ones(D), mp, sp, n_jobs)


def free_energy_of(setup, n_jobs=None):
    """Instantiate the reference's FreeEnergy for a setup."""
    C = harness.reference_classes()
    return C["FreeEnergy"](setup["sde"], setup["mu0"], setup["tau0"],
                           setup["obs_times"], setup["obs_values"],
                           setup["obs_noise"], setup["obs_mask"],
                           n_jobs=setup["n_jobs"] if n_jobs is None else n_jobs)
