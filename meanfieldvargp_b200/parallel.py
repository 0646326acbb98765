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


def owned_columns(M, n0, n1):
    """
    Columns of the mean block [D][3M+4] and of the variance block [D][2M+3]
    whose entries of x and of the gradient rank ``[n0, n1)`` owns.  The column
    shared by two neighbouring blocks (3*n1 / 2*n1) belongs to the right-hand
    block; the last block also owns the columns the reference never integrates
    (free_energy.py:168, 490, 719).
    """
    L = M - 1
    last = n1 == L
    cm = np.arange(3 * n0, (3 * M + 4) if last else 3 * n1)
    cs = np.arange(2 * n0, (2 * M + 3) if last else 2 * n1)
    return cm, cs


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
