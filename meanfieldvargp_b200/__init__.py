"""
meanfieldvargp_b200 -- B200-native (sm_100a, FP64 CUDA) free-energy path of
vrettasm/MeanFieldVarGP: E_sde + gradients, E_kl0, E_obs, E_cost and the SCG
loop, behind the reference's own FreeEnergy / SCG Python API and the C ABI of
include/mfvgp.h.
"""
from ._lib import MfvgpError, load as load_library          # noqa: F401
from .free_energy import FreeEnergy                          # noqa: F401
from .scaled_cg import SCG                                   # noqa: F401
from .posterior import (initial_path, initial_state, pack,   # noqa: F401
                        reconstruct, support_indices, time_knots, unpack)
from .trajectory import (collect_obs, simulate_dw, simulate_l63,   # noqa: F401
                         simulate_l96, simulate_ou)
from .systems import (DoubleWell, Lorenz63, Lorenz96,        # noqa: F401
                      OrnsteinUhlenbeck, StochasticProcess)

__all__ = ["FreeEnergy", "SCG", "DoubleWell", "OrnsteinUhlenbeck", "Lorenz63", "Lorenz96",
           "StochasticProcess", "MfvgpError", "load_library", "reconstruct",
           "support_indices", "pack", "unpack", "time_knots", "initial_path", "initial_state", "simulate_l96", "simulate_dw", "simulate_ou",
           "simulate_l63", "collect_obs"]
