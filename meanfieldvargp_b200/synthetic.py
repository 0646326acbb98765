"""
Synthetic Lorenz 96 problems of SURVEY.md 8(d) (no trajectory simulation): what bench.py
measures on and what the size-independent tests use.  Seeded; FP64.

The north-star configuration is ``synthetic_l96(500, 16, seed=8192)`` (L96 D=500, d=125, M=16:
the shape of examples/example_500D.ipynb), the long-horizon one
``synthetic_l96_device(torch, 8192, 4097, seed=8192)`` (state generated on the GPU).
"""
import numpy as np


def synthetic_l96(D, M, seed, dt_obs=0.25, partial=True):
    """dict of FreeEnergy constructor inputs + x0 (host arrays)"""
    rng = np.random.default_rng(seed)
    Nm, Ns = 3 * M + 4, 2 * M + 3
    obs_times = dt_obs * np.arange(1, M + 1)
    raw = rng.normal(2.3, 3.6, size=(D, Nm + 3))
    mp = (raw[:, :-3] + raw[:, 1:-2] + raw[:, 2:-1] + raw[:, 3:]) / 4.0
    sp = np.log(4.0 + 2.0 * rng.random((D, Ns)))
    mask = [bool(i % 4 == 0) for i in range(D)] if partial else [True] * D
    d = sum(mask)
    obs_values = rng.normal(2.3, 3.6, size=(M, d)) + np.sqrt(2.0) * rng.standard_normal((M, d))
    return dict(system="L96", sigma=40.0 * np.ones(D), theta=np.array([8.0]),
                mu0=10.0 * np.ones(D), tau0=4.0 * np.ones(D), obs_times=obs_times,
                obs_values=obs_values, obs_noise=2.0 * np.ones(d), obs_mask=mask,
                x0=np.concatenate((mp.ravel(), sp.ravel())))


def synthetic_l96_device(torch, D, M, seed, dt_obs=0.25):
    """Same recipe with x0 generated on the GPU (large D: 1.3 GB at D=8192, M=4097)."""
    gen = torch.Generator(device="cuda")
    gen.manual_seed(seed)
    Nm, Ns = 3 * M + 4, 2 * M + 3
    raw = 2.3 + 3.6 * torch.randn((D, Nm + 3), dtype=torch.float64, device="cuda",
                                  generator=gen)
    mp = (raw[:, :-3] + raw[:, 1:-2] + raw[:, 2:-1] + raw[:, 3:]) / 4.0
    sp = torch.log(4.0 + 2.0 * torch.rand((D, Ns), dtype=torch.float64, device="cuda",
                                          generator=gen))
    x0 = torch.cat((mp.reshape(-1), sp.reshape(-1))).contiguous()
    rng = np.random.default_rng(seed)
    mask = [bool(i % 4 == 0) for i in range(D)]
    d = sum(mask)
    obs_values = rng.normal(2.3, 3.6, size=(M, d)) + np.sqrt(2.0) * rng.standard_normal((M, d))
    return dict(system="L96", sigma=40.0 * np.ones(D), theta=np.array([8.0]),
                mu0=10.0 * np.ones(D), tau0=4.0 * np.ones(D),
                obs_times=dt_obs * np.arange(1, M + 1), obs_values=obs_values,
                obs_noise=2.0 * np.ones(d), obs_mask=mask, x0=x0)


def make_sde(g):
    """the parameter holder (systems.py) of a problem dict"""
    from . import systems
    sysn = g["system"]
    if sysn == "DW":
        sde = systems.DoubleWell(g["sigma"][0], g["theta"][0])
    elif sysn == "OU":
        sde = systems.OrnsteinUhlenbeck(g["sigma"][0], g["theta"])
    elif sysn == "L63":
        sde = systems.Lorenz63(g["sigma"], g["theta"])
    else:
        sde = systems.Lorenz96(g["sigma"], g["theta"][0], dim_D=np.asarray(g["sigma"]).size)
    sde.time_window = np.array([0.0, float(g["obs_times"][-1]) + 1.0])
    return sde


def make_free_energy(g, **kw):
    from .free_energy import FreeEnergy
    return FreeEnergy(make_sde(g), g["mu0"], g["tau0"], g["obs_times"], g["obs_values"],
                      g["obs_noise"], g["obs_mask"], **kw)
