"""
Multi-GPU evaluation: the L = M-1 time intervals are sharded over the ranks
(one process per GPU), as the reference itself treats them as independent
tasks (free_energy.py:500-508).  See DESIGN.md "Multi-GPU".
"""
import numpy as np

from .free_energy import FreeEnergy


def shard_intervals(L, rank, world):
    """Contiguous block of intervals for ``rank`` (balanced to within one)."""
    base, rem = divmod(L, world)
    n0 = rank * base + min(rank, rem)
    return n0, n0 + base + (1 if rank < rem else 0)


class ShardedFreeEnergy(object):
    """E_cost of one problem evaluated by ``world`` ranks, intervals sharded."""

    def __init__(self, g, rank=0, world=1, device=None):
        import torch
        from . import systems
        self.rank, self.world = rank, world
        D = int(np.asarray(g["sigma"]).size)
        sde = systems.Lorenz96(g["sigma"], float(np.atleast_1d(g["theta"])[0]), dim_D=D)
        sde.time_window = np.array([0.0, float(g["obs_times"][-1]) + 1.0])
        L = len(g["obs_times"]) - 1
        if world > L:
            raise ValueError(f"cannot shard {L} intervals over {world} ranks")
        self.range = shard_intervals(L, rank, world)
        dev = torch.cuda.current_device() if device is None else device
        self.local = FreeEnergy(sde, g["mu0"], g["tau0"], g["obs_times"], g["obs_values"],
                                g["obs_noise"], g["obs_mask"], device=dev,
                                intervals=self.range if world > 1 else None)
        if world > 1:
            self.local.attach_communicator(rank, world)

    def E_cost(self, x, output_gradients=True):
        return self.local.E_cost(x, output_gradients)
