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
    """
    E_cost / find_minimum / grad_params of one problem evaluated by ``world`` ranks, the L = M-1
    time intervals sharded in contiguous blocks (any of the four systems).  ``g``: the
    constructor inputs as a dict (system, sigma, theta, mu0, tau0, obs_times, obs_values,
    obs_noise, obs_mask).  A rank needs at least one interval: world <= L (SURVEY.md 8(e)'s
    alternative for L < world, sharding the state dimensions with a 3-row halo, moves three
    rows of every column per evaluation and is not built: such problems are far below one
    wave of one GPU).
    """

    def __init__(self, g, rank=0, world=1, device=None):
        import torch
        from .synthetic import make_sde
        self.rank, self.world = rank, world
        g = dict(g)
        g.setdefault("system", "L96")
        g["theta"] = np.atleast_1d(np.asarray(g["theta"], dtype=float))
        g["sigma"] = np.atleast_1d(np.asarray(g["sigma"], dtype=float))
        sde = make_sde(g)
        L = len(g["obs_times"]) - 1
        if world > L:
            raise ValueError(f"cannot shard {L} intervals over {world} ranks")
        self.range = shard_intervals(L, rank, world)
        dev = torch.cuda.current_device() if device is None else device
        self.local = FreeEnergy(sde, g["mu0"], g["tau0"], g["obs_times"], g["obs_values"],
                                g["obs_noise"], g.get("obs_mask"), device=dev,
                                intervals=self.range if world > 1 else None)
        if world > 1:
            self.local.attach_communicator(rank, world)

    def E_cost(self, x, output_gradients=True):
        return self.local.E_cost(x, output_gradients)

    def find_minimum(self, x0, **kw):
        return self.local.find_minimum(x0, **kw)

    def grad_params(self, x):
        return self.local.grad_params(x)
