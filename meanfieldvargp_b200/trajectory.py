"""
Lorenz 96 sample paths and observations on the device: the
reference's ``Lorenz96.make_trajectory`` (src/dynamical_systems/lorenz_96.py:103-185) and
``StochasticProcess.collect_obs`` (stochastic_process.py:211-297) -- the step that produces
the data every notebook starts from, and what generates *physical* long-horizon test data
for D = 8192 (the burn-in alone is 125 D steps).

Lorenz 96 is chaotic, so the CUDA kernel (csrc/trajectory.cuh) reproduces the reference's
arithmetic bit for bit; the noise is drawn exactly as the reference draws it
(``default_rng(r_seed).standard_normal((D, dim_t))``, all up front) unless given.
No CPU fallback: needs the library and a GPU.
"""
import ctypes

import numpy as np

from . import _lib


def time_window(t0, tf, dt):
    """``np.arange(t0, tf + dt, dt)`` (lorenz_96.py:118)."""
    return np.arange(t0, tf + dt, dt, dtype=float)


def simulate_l96(sigma, theta, dim_D, t0, tf, dt=0.01, r_seed=None, noise=None):
    """
    Sample path of the stochastic Lorenz 96 system.

    :param sigma: diffusion coefficient (scalar or [D], lorenz_96.py:70-91)
    :param theta: forcing constant
    :param noise: optional standard-normal array [D, dim_t] (the reference's layout);
                  default: ``np.random.default_rng(r_seed).standard_normal((D, dim_t))``
    :return: ``(tk [dim_t], x [dim_t, D])`` as CUDA ``torch.float64`` tensors -- the
             reference's ``time_window`` and ``sample_path``
    """
    import torch
    D = int(dim_D)
    if D < 4:
        raise ValueError(f" Lorenz96: Insufficient state vector dimensions: {D}")
    sigma = np.asarray(sigma, dtype=float)
    sigma = sigma * np.ones(D) if sigma.ndim == 0 else sigma
    if sigma.shape != (D,):
        raise ValueError(f" Lorenz96: Wrong matrix dimensions: {sigma.shape}")
    tk = time_window(t0, tf, dt)
    dim_t = tk.size
    lib = _lib.require_device()
    if noise is None:
        noise = np.random.default_rng(r_seed).standard_normal((D, dim_t))
    if type(noise).__module__.startswith("torch"):
        z = noise.to(device="cuda", dtype=torch.float64)
    else:
        z = torch.as_tensor(np.asarray(noise, dtype=float), device="cuda")
    if tuple(z.shape) != (D, dim_t):
        raise ValueError(f"simulate_l96: noise must be [{D}, {dim_t}]")
    # ek = sqrt(sigma dt) * N(0,1), then time major (lorenz_96.py:153-162, :177 uses ek.T[t])
    ek = (torch.as_tensor(np.sqrt(sigma * dt), device="cuda")[:, None] * z).T.contiguous()
    x = torch.empty((dim_t, D), dtype=torch.float64, device="cuda")
    status = ctypes.c_int32()
    stream = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.mfvgp_l96_trajectory(D, float(theta), float(dt), dim_t, 125 * D, 1.0e-3,
                                        ek.data_ptr(), x.data_ptr(), ctypes.byref(status), stream),
               "mfvgp_l96_trajectory")
    if status.value & 1:                                             # lorenz_96.py:142-145
        raise RuntimeError(" Lorenz96: Invalid initial point (NaN after the burn-in).")
    if status.value & 2:                                             # lorenz_96.py:170-174
        raise RuntimeError(f" Lorenz96: Invalid state vector. Reduce the value of dt={dt} "
                           f"and try again.")
    return torch.as_tensor(tk, device="cuda"), x


def collect_obs(tk, xt, n_obs, h_mask=None):
    """
    Noise-free observations from a sample path (stochastic_process.py:211-297): ``n_obs``
    observations per time unit (int) or explicit time indices (list).  Returns
    ``(obs_idx, obs_values)``; works on numpy arrays and on tensors.
    """
    dim_t = len(tk)
    if isinstance(n_obs, int):
        dt = float(abs(tk[1] - tk[0]))
        if n_obs > int(1.0 / dt):                                    # :253-255
            raise ValueError(" Observation density exceeds the number of samples.")
        dim_m = int(np.floor(abs(float(tk[0]) - float(tk[-1])) * n_obs))      # :259
        idx = np.linspace(0, dim_t, dim_m + 2, dtype=int)            # :265
        obs_t = [int(i) for i in np.sort(np.unique(idx[1:-1]))]      # :268
    elif isinstance(n_obs, list):
        obs_t = sorted(np.unique(n_obs))                             # :273
        if len(obs_t) > dim_t:
            raise ValueError(" Observation density exceeds the number of samples.")
        obs_t = list(map(int, obs_t))
    else:
        raise TypeError("collect_obs: n_obs must be an int or a list of indices")
    obs_y = xt[obs_t]                                                # :286
    if h_mask is not None:                                           # :289-293
        obs_y = obs_y[:, [i for i, m in enumerate(h_mask) if m]]
    return obs_t, obs_y
