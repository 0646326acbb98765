"""
TEST INFRASTRUCTURE -- CPU restatement (numpy) of Lorenz96.make_trajectory
(src/dynamical_systems/lorenz_96.py:9-30, :103-185), the checker for csrc/trajectory.cuh.
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
