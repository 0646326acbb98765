"""
Sample paths and observations on the device: the
reference's ``Lorenz96.make_trajectory`` (src/dynamical_systems/lorenz_96.py:103-185) and
``StochasticProcess.collect_obs`` (stochastic_process.py:211-297) -- the step that produces
the data every notebook starts from, and what generates *physical* long-horizon test data
for D = 8192 (the burn-in alone is 125 D steps).

Lorenz 96 is chaotic, so the CUDA kernel (csrc/trajectory.cuh) reproduces the reference's
arithmetic bit for bit; the noise is drawn exactly as the reference draws it
(``default_rng(r_seed).standard_normal((D, dim_t))``, all up front) unless given.
The one- and three-dimensional systems (DoubleWell / OrnsteinUhlenbeck / Lorenz63
``make_trajectory``: double_well.py:40-91, ornstein_uhlenbeck.py:38-79,
lorenz_63.py:102-181) follow the same rule; a path there is one thread's work, and the
device's contribution is the ensemble: ``npaths`` independent paths per call.
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


_SMALL = {"DW": (0, 1, 1, 0), "OU": (1, 2, 1, 0), "L63": (2, 3, 3, 5000)}   # id, ntheta, dims, burn-in


def _small_paths(system, theta, dt, x0, ek):
    """x0 [npaths, dims], ek [npaths, dim_t, dims] (CUDA float64) -> x [npaths, dim_t, dims]."""
    import torch
    sys_id, ntheta, dims, nburn = _SMALL[system]
    lib = _lib.require_device()
    theta = np.ascontiguousarray(np.atleast_1d(np.asarray(theta, dtype=float)))
    if theta.size != ntheta:
        raise ValueError(f"{system}: expected {ntheta} drift parameter(s), got {theta.size}")
    npaths, dim_t = int(ek.shape[0]), int(ek.shape[1])
    x = torch.empty((npaths, dim_t, dims), dtype=torch.float64, device="cuda")
    status = np.zeros(npaths, dtype=np.int32)
    stream = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.mfvgp_small_trajectory(sys_id, theta.ctypes.data, ntheta, float(dt), dim_t, nburn,
                                          1.0e-3, npaths, x0.data_ptr(), ek.data_ptr(),
                                          x.data_ptr(), status.ctypes.data, stream),
               "mfvgp_small_trajectory")
    return x, status


def _rngs(r_seed, rng, npaths):
    if rng is not None:
        if npaths != 1:
            raise ValueError("an explicit generator drives one path; use r_seed for an ensemble")
        return [rng]
    if npaths == 1:
        return [np.random.default_rng(r_seed)]
    return [np.random.default_rng(s) for s in np.random.SeedSequence(r_seed).spawn(npaths)]


def simulate_dw(sigma, theta, t0, tf, dt=0.01, r_seed=None, rng=None, npaths=1):
    """
    Double-well sample path(s), ``DoubleWell.make_trajectory`` (double_well.py:40-91): start at
    +-theta (sign flipped when ``rng.random() > 0.5``) plus ``sqrt(sigma dt / 2) N(0,1)``, then
    Euler-Maruyama.  The generator is consumed in the reference's order, so ``npaths=1`` with
    the reference's seed gives the reference's path.  Returns ``(tk, x)``: x ``[dim_t]`` for one
    path, ``[npaths, dim_t]`` for an ensemble (path p uses the p-th child of
    ``SeedSequence(r_seed)``); CUDA float64 tensors.
    """
    import torch
    tk = time_window(t0, tf, dt)
    dim_t = tk.size
    sigma, theta = float(np.asarray(sigma).reshape(-1)[0]), float(np.asarray(theta).reshape(-1)[0])
    x0 = np.empty((npaths, 1))
    ek = np.empty((npaths, dim_t, 1))
    for p, g in enumerate(_rngs(r_seed, rng, npaths)):
        v = theta
        if g.random() > 0.5:                                         # :69-71
            v *= -1.0
        x0[p, 0] = v + np.sqrt(0.5 * sigma * dt) * g.standard_normal()      # :74
        ek[p, :, 0] = np.sqrt(sigma * dt) * g.standard_normal(dim_t)        # :77
    x, _ = _small_paths("DW", [theta], dt, torch.as_tensor(x0, device="cuda"),
                        torch.as_tensor(ek, device="cuda"))
    x = x[:, :, 0]
    return torch.as_tensor(tk, device="cuda"), (x[0] if npaths == 1 else x)


def simulate_ou(sigma, theta, t0, tf, dt=0.01, r_seed=None, rng=None, npaths=1):
    """
    Ornstein-Uhlenbeck sample path(s), ``OrnsteinUhlenbeck.make_trajectory``
    (ornstein_uhlenbeck.py:38-79): ``theta = (theta, mu)``, start at mu.  Shapes as ``simulate_dw``.
    """
    import torch
    tk = time_window(t0, tf, dt)
    dim_t = tk.size
    sigma = float(np.asarray(sigma).reshape(-1)[0])
    th, mu = (float(v) for v in np.asarray(theta, dtype=float).reshape(-1))      # :53
    x0 = np.full((npaths, 1), mu)                                    # :65
    ek = np.empty((npaths, dim_t, 1))
    for p, g in enumerate(_rngs(r_seed, rng, npaths)):
        ek[p, :, 0] = np.sqrt(sigma * dt) * g.standard_normal(dim_t)             # :68
    x, _ = _small_paths("OU", [th, mu], dt, torch.as_tensor(x0, device="cuda"),
                        torch.as_tensor(ek, device="cuda"))
    x = x[:, :, 0]
    return torch.as_tensor(tk, device="cuda"), (x[0] if npaths == 1 else x)


def simulate_l63(sigma, theta, t0, tf, dt=0.01, r_seed=None, rng=None, npaths=1):
    """
    Lorenz 63 sample path(s), ``Lorenz63.make_trajectory`` (lorenz_63.py:102-181): burn-in of
    5000 Euler steps of 1e-3 from (1, 1, 1), then Euler-Maruyama with
    ``sqrt(sigma dt) * standard_normal((3, dim_t))``.  x ``[dim_t, 3]`` for one path,
    ``[npaths, dim_t, 3]`` for an ensemble.
    """
    import torch
    tk = time_window(t0, tf, dt)
    dim_t = tk.size
    sigma = np.asarray(sigma, dtype=float)
    sigma = sigma * np.ones(3) if sigma.ndim == 0 else (sigma.diagonal() if sigma.ndim == 2 else sigma)
    theta = np.asarray(theta, dtype=float).reshape(-1)
    x0 = np.ones((npaths, 3))                                        # :124
    ek = np.empty((npaths, dim_t, 3))
    scale = np.repeat(np.sqrt(sigma * dt), dim_t).reshape(3, dim_t)  # :150-153
    for p, g in enumerate(_rngs(r_seed, rng, npaths)):
        ek[p] = (scale * g.standard_normal((3, dim_t))).T            # :156, :172
    x, status = _small_paths("L63", theta, dt, torch.as_tensor(x0, device="cuda"),
                             torch.as_tensor(ek, device="cuda"))
    if np.any(status & 1):                                           # :136-139
        raise RuntimeError(" Lorenz63: Invalid initial point (NaN after the burn-in).")
    if np.any(status & 2):                                           # :164-168
        raise RuntimeError(f" Lorenz63: Invalid state vector. Reduce the value of dt={dt} "
                           f"and try again.")
    return torch.as_tensor(tk, device="cuda"), (x[0] if npaths == 1 else x)


def collect_obs(tk, xt, n_obs, h_mask=None, owner="collect_obs"):
    """
    Noise-free observations from a sample path (stochastic_process.py:211-297): ``n_obs``
    observations per time unit (int) or explicit time indices (list).  Returns
    ``(obs_idx, obs_values)``; works on numpy arrays and on tensors.
    """
    dim_t = len(tk)
    if isinstance(n_obs, int):
        dt = float(abs(tk[1] - tk[0]))
        if n_obs > int(1.0 / dt):                                    # :253-255
            raise ValueError(f" {owner}: Observation density exceeds the number of samples.")
        dim_m = int(np.floor(abs(float(tk[0]) - float(tk[-1])) * n_obs))      # :259
        idx = np.linspace(0, dim_t, dim_m + 2, dtype=int)            # :265
        obs_t = [int(i) for i in np.sort(np.unique(idx[1:-1]))]      # :268
    elif isinstance(n_obs, list):
        obs_t = sorted(np.unique(n_obs))                             # :273
        if len(obs_t) > dim_t:
            raise ValueError(f" {owner}: Observation density exceeds the number of samples.")
        obs_t = list(map(int, obs_t))
    else:
        raise TypeError("collect_obs: n_obs must be an int or a list of indices")
    obs_y = xt[obs_t]                                                # :286
    if h_mask is not None:                                           # :289-293
        obs_y = obs_y[:, [i for i, m in enumerate(h_mask) if m]]
    return obs_t, obs_y
