"""Shared helpers for the parity tests (golden fixtures, problem builders)."""
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"

EVAL_CASES = ["DW", "OU", "L63", "L96_4", "L96_8", "L96_40", "L96_67p", "L96_500p",
              "L96_5_irregular"]
REFINE_CASES = ["DW_refine", "L96_6_refine"]
SCG_CASES = ["DW", "OU", "L63_T5", "L96_8"]
# reference runs that take the rare branches of scaled_cg.py: rejected trial points (:254-279),
# mu >= 0 (:194-197), delta <= 0 (:233-236), restart after dim_x successes (:361-364)
SCG_BRANCH_CASES = ["branch_DW3", "branch_DW4", "hilltop_DW3", "hilltop_DW3b"]


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


from meanfieldvargp_b200.synthetic import (make_free_energy, make_sde,  # noqa: E402,F401
                                           synthetic_l96, synthetic_l96_device)


def make_oracle(g, adaptive=True):
    from oracle import mfvgp_oracle as O
    return O.FreeEnergyOracle(g["system"], g["sigma"], g["theta"], g["mu0"], g["tau0"],
                              g["obs_times"], g["obs_values"], g["obs_noise"], g["obs_mask"],
                              adaptive=adaptive)
