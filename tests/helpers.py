"""Shared helpers for the parity tests (golden fixtures, problem builders)."""
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"

EVAL_CASES = ["DW", "OU", "L63", "L96_4", "L96_8", "L96_40", "L96_67p", "L96_500p",
              "L96_5_irregular"]
REFINE_CASES = ["DW_refine", "L96_6_refine"]
SCG_CASES = ["DW", "OU", "L63_T5", "L96_8"]
# reference runs that take the rare branches of scaled_cg.py: rejected trial points (:254-279),
# mu >= 0 (:194-197), restart after dim_x successes (:361-364)
SCG_BRANCH_CASES = ["branch_DW3", "branch_DW4"]


def load_golden(kind, name):
    z = np.load(GOLDEN / f"{kind}_{name}.npz")
    g = {k: z[k] for k in z.files}
    g["system"] = str(g["system"])
    g["obs_mask"] = None if g["obs_mask"].size == 0 else [bool(b) for b in g["obs_mask"]]
    return g


def rel_err(a, b):
    """max |a-b| / max |b| (the 1e-10 bar of BASELINE.json is on this)."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    den = max(float(np.abs(b).max()), 1e-300)
    return float(np.abs(a - b).max()) / den


def worst_elem_err(a, b, floor=1.0e-3):
    """
    Worst element-wise relative error: max_i |a_i - b_i| / max(|b_i|, floor * max|b|).
    rel_err above is a max-norm and never looks at small entries on their own; this does, down
    to entries `floor` times the largest (below that an entry is the rounding residue of
    cancelling terms of the size of the largest one, so 1e-16 / floor is what FP64 can give).
    """
    a, b = np.asarray(a, dtype=float).ravel(), np.asarray(b, dtype=float).ravel()
    den = np.maximum(np.abs(b), floor * max(float(np.abs(b).max()), 1e-300))
    return float((np.abs(a - b) / den).max())


def make_sde(g):
    """Our parameter holder for a golden case."""
    import meanfieldvargp_b200 as mf
    sysn = g["system"]
    if sysn == "DW":
        sde = mf.DoubleWell(g["sigma"][0], g["theta"][0])
    elif sysn == "OU":
        sde = mf.OrnsteinUhlenbeck(g["sigma"][0], g["theta"])
    elif sysn == "L63":
        sde = mf.Lorenz63(g["sigma"], g["theta"])
    else:
        sde = mf.Lorenz96(g["sigma"], g["theta"][0], dim_D=g["sigma"].size)
    sde.time_window = np.array([0.0, float(g["obs_times"][-1]) + 1.0])
    return sde


def make_free_energy(g, **kw):
    import meanfieldvargp_b200 as mf
    return mf.FreeEnergy(make_sde(g), g["mu0"], g["tau0"], g["obs_times"], g["obs_values"],
                         g["obs_noise"], g["obs_mask"], **kw)


def make_oracle(g, adaptive=True):
    from oracle import mfvgp_oracle as O
    return O.FreeEnergyOracle(g["system"], g["sigma"], g["theta"], g["mu0"], g["tau0"],
                              g["obs_times"], g["obs_values"], g["obs_noise"], g["obs_mask"],
                              adaptive=adaptive)


def synthetic_l96(D, M, seed, dt_obs=0.25, partial=True):
    """SURVEY.md 8(d) synthetic L96 state (no trajectory simulation)."""
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
    """Same recipe as synthetic_l96 but x0 is generated on the GPU (large D)."""
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
