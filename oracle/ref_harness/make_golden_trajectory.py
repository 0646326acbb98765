"""
Generates tests/golden/trajectory_*.npz by running the UNMODIFIED reference:
``Lorenz96(sigma, theta, dim_D, r_seed).make_trajectory(t0, tf, dt)``
(src/dynamical_systems/lorenz_96.py:103-185) and ``collect_obs``
(stochastic_process.py:211-297); likewise DoubleWell / OrnsteinUhlenbeck / Lorenz63
(double_well.py:40-91, ornstein_uhlenbeck.py:38-79, lorenz_63.py:102-181).
Run in the build container:

    python oracle/ref_harness/make_golden_trajectory.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.ref_harness import harness  # noqa: E402

CASES = {
    # name: (D, sigma, theta, t0, tf, dt, seed, n_obs)
    "D4": (4, 10.0, 8.0, 0.0, 0.5, 0.01, 575, 4),
    "D40": (40, 40.0, 8.0, 0.0, 0.6, 0.001, 575, 10),
    "D67": (67, 10.0, 8.0, 0.0, 0.5, 0.002, 7, [10, 50, 120]),
}

SMALL_CASES = {
    # name: (class, sigma, theta, t0, tf, dt, seed, n_obs)
    "DW": ("DoubleWell", 0.8, 1.0, 0.0, 10.0, 0.001, 575, 2),           # example_DW.ipynb
    "DW_coarse": ("DoubleWell", 1.2, 0.7, 1.0, 6.0, 0.01, 3, 5),
    "OU": ("OrnsteinUhlenbeck", 0.5, [2.0, 1.5], 0.0, 5.0, 0.01, 11, 3),
    "L63": ("Lorenz63", [40.0, 40.0, 40.0], [10.0, 28.0, 2.66667], 0.0, 4.0, 0.001, 575, 3),
    "L63_scalar_sigma": ("Lorenz63", 10.0, [10.0, 28.0, 8.0 / 3.0], 0.0, 2.0, 0.005, 21, 4),
}


def main():
    cls = harness.reference_classes()
    out = ROOT / "tests" / "golden"
    for name, (D, sigma, theta, t0, tf, dt, seed, n_obs) in CASES.items():
        sde = cls["Lorenz96"](sigma, theta, dim_D=D, r_seed=seed)
        sde.make_trajectory(t0, tf, dt)
        mask = [bool(i % 3 == 0) for i in range(D)]
        obs_t, obs_y = sde.collect_obs(n_obs, mask)
        np.savez_compressed(out / f"trajectory_{name}.npz", D=D, sigma=sigma, theta=theta, t0=t0,
                            tf=tf, dt=dt, seed=seed, tk=sde.time_window, x=sde.sample_path,
                            n_obs=np.asarray(n_obs), obs_t=np.asarray(obs_t), obs_y=obs_y,
                            mask=np.asarray(mask))
        print(name, sde.sample_path.shape, len(obs_t), obs_y.shape)
    for name, (cname, sigma, theta, t0, tf, dt, seed, n_obs) in SMALL_CASES.items():
        sde = cls[cname](sigma, theta, r_seed=seed)
        sde.make_trajectory(t0, tf, dt)
        obs_t, obs_y = sde.collect_obs(n_obs)
        np.savez_compressed(out / f"trajectory_{name}.npz", cls=cname, sigma=np.asarray(sigma),
                            theta=np.asarray(theta), t0=t0, tf=tf, dt=dt, seed=seed,
                            tk=sde.time_window, x=sde.sample_path, n_obs=np.asarray(n_obs),
                            obs_t=np.asarray(obs_t), obs_y=obs_y)
        print(name, sde.sample_path.shape, len(obs_t), obs_y.shape)


if __name__ == "__main__":
    main()
