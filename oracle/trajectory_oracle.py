"""
TEST INFRASTRUCTURE -- CPU restatement (numpy) of Lorenz96.make_trajectory
(src/dynamical_systems/lorenz_96.py:9-30, :103-185) and of the DoubleWell /
OrnsteinUhlenbeck / Lorenz63 ones (double_well.py:40-91, ornstein_uhlenbeck.py:38-79,
lorenz_63.py:9-35, :102-181), the checker for csrc/trajectory.cuh.
Only tests/, smoke() and bench.py's CPU legs may import this.

Same operations in the same order as the reference (np.roll drift, explicit Euler burn-in,
Euler-Maruyama with all the noise drawn up front), so the result is bit-identical to the
reference's -- pinned by tests/golden/trajectory_*.npz, which the unmodified reference
produced (oracle/ref_harness/make_golden_trajectory.py).
"""
import numpy as np


def l96(x, u):
    """lorenz_96.py:9-30."""
    return (np.roll(x, -1) - np.roll(x, +2)) * np.roll(x, +1) - x + u


def make_trajectory(sigma, theta, dim_D, t0, tf, dt=0.01, r_seed=None, noise=None):
    """(tk, sample_path [dim_t, D]) -- lorenz_96.py:103-185."""
    D = int(dim_D)
    sigma = np.asarray(sigma, dtype=float)
    sigma = sigma * np.ones(D) if sigma.ndim == 0 else sigma
    tk = np.arange(t0, tf + dt, dt, dtype=float)
    dim_t = tk.size
    x0 = theta * np.ones(D)
    delta_t = 1.0e-3
    x0[int(D / 2.0)] += 0.01
    for _ in range(125 * D):
        x0 = x0 + l96(x0, theta) * delta_t
    x = np.zeros((dim_t, D), dtype=float)
    x[0] = x0
    ek = np.sqrt(sigma * dt)
    ek = np.repeat(ek, dim_t).reshape(D, dim_t)
    if noise is None:
        noise = np.random.default_rng(r_seed).standard_normal((D, dim_t))
    ek = ek * noise
    for t in range(1, dim_t):
        x[t] = x[t - 1] + l96(x[t - 1], theta) * dt + ek.T[t]
    return tk, x


def make_trajectory_dw(sigma, theta, t0, tf, dt=0.01, r_seed=None, rng=None):
    """(tk, x [dim_t]) -- double_well.py:40-91."""
    rng = np.random.default_rng(r_seed) if rng is None else rng
    tk = np.arange(t0, tf + dt, dt, dtype=float)
    dim_t = tk.size
    x = np.zeros(dim_t)
    x[0] = theta
    if rng.random() > 0.5:
        x[0] *= -1.0
    x[0] += np.sqrt(0.5 * sigma * dt) * rng.standard_normal()
    ek = np.sqrt(sigma * dt) * rng.standard_normal(dim_t)
    for t in range(1, dim_t):
        x[t] = x[t - 1] + 4.0 * x[t - 1] * (theta - x[t - 1] ** 2) * dt + ek[t]
    return tk, x


def make_trajectory_ou(sigma, theta, t0, tf, dt=0.01, r_seed=None, rng=None):
    """(tk, x [dim_t]) -- ornstein_uhlenbeck.py:38-79; theta = (theta, mu)."""
    rng = np.random.default_rng(r_seed) if rng is None else rng
    th, mu = theta
    tk = np.arange(t0, tf + dt, dt, dtype=float)
    dim_t = tk.size
    x = np.zeros(dim_t)
    x[0] = mu
    ek = np.sqrt(sigma * dt) * rng.standard_normal(dim_t)
    for t in range(1, dim_t):
        x[t] = x[t - 1] + th * (mu - x[t - 1]) * dt + ek[t]
    return tk, x


def l63(state, theta):
    """lorenz_63.py:9-35."""
    x, y, z = state
    sigma, rho, beta = theta
    return np.array([sigma * (y - x), (rho - z) * x - y, x * y - beta * z])


def make_trajectory_l63(sigma, theta, t0, tf, dt=0.01, r_seed=None, rng=None):
    """(tk, x [dim_t, 3]) -- lorenz_63.py:102-181."""
    rng = np.random.default_rng(r_seed) if rng is None else rng
    sigma = np.asarray(sigma, dtype=float) * np.ones(3)
    theta = np.asarray(theta, dtype=float)
    tk = np.arange(t0, tf + dt, dt, dtype=float)
    dim_t = tk.size
    x0 = np.ones(3)
    for _ in range(5000):
        x0 = x0 + l63(x0, theta) * 1.0e-3
    x = np.zeros((dim_t, 3))
    x[0] = x0
    ek = np.repeat(np.sqrt(sigma * dt), dim_t).reshape(3, dim_t)
    ek = ek * rng.standard_normal((3, dim_t))
    for t in range(1, dim_t):
        x[t] = x[t - 1] + l63(x[t - 1], theta) * dt + ek.T[t]
    return tk, x
