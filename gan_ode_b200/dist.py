"""Data parallelism for the latent-motion path: one process per GPU, trajectories shard, weights are
replicated, and the only exchange is one all-reduce (sum) of the flat ODE parameter-gradient buffer the backward
kernel wrote, issued on the same stream right behind it (NCCL over NVLink on a B200 box; gloo in the CPU tests).
There is no data-path collective: trajectories are independent.  dopri5's error norm stays per rank, i.e. what N
independent reference processes would compute."""
from __future__ import annotations

import importlib
from typing import Tuple

import torch
import torch.distributed as dist

_api = importlib.import_module(__package__ + ".odeint")

__all__ = ["shard_bounds", "shard_batch", "enable_grad_allreduce", "disable_grad_allreduce", "philox_for_shard"]


def shard_bounds(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced shard [lo, hi) of a batch of B trajectories (first B % world ranks get one extra)."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(y0: torch.Tensor, rank: int = None, world: int = None):
    """This rank's rows of y0 and the global index of its first trajectory (the Philox `traj_offset`)."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(y0.shape[0], rank, world)
    return y0[lo:hi], lo


def enable_grad_allreduce(group=None):
    """After this, every backward of odeint / odeint_adjoint / sdeint all-reduces its flat parameter gradient in place
    (sum over ranks) before handing the views to autograd."""
    _api.config.grad_allreduce = True if group is None else group


def disable_grad_allreduce():
    _api.config.grad_allreduce = None


def philox_for_shard(seed: int, y0: torch.Tensor, rank: int = None, world: int = None):
    """(local y0, PhiloxBrownian keyed by global trajectory index) so the SDE sampler draws the same increments for a
    trajectory no matter which rank integrates it."""
    from .sdeint import PhiloxBrownian
    local, lo = shard_batch(y0, rank, world)
    return local, PhiloxBrownian(seed, traj_offset=lo)
