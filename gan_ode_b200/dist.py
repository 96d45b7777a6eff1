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

__all__ = ["shard_bounds", "shard_batch", "enable_grad_allreduce", "disable_grad_allreduce", "philox_for_shard",
           "enable_p2p_allreduce", "P2PAllReduce"]


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


class P2PAllReduce:
    """One-shot all-reduce of small fp32 vectors over NVLink peer memory (csrc/p2p_allreduce.cu), built on torch's
    symmetric-memory rendezvous for the address exchange only.  `cap` floats per vector at most."""

    small = 8192  # floats up to which the single-CTA kernel beats ncclAllReduce (measured: 544 floats 11 vs 17 us at N=2)

    def __init__(self, group=None, cap: int = 65536):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        group = dist.group.WORLD if group is None else group
        self.world, self.rank, self.cap = dist.get_world_size(group), dist.get_rank(group), int(cap)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.buf = symm_mem.empty(2 * self.world * self.cap, dtype=torch.float32, device=dev)
        self.hdl = symm_mem.rendezvous(self.buf, group)
        if self.hdl.signal_pad_size < 8 * self.world:
            raise RuntimeError("symmetric-memory signal pad too small")
        self.epoch = torch.zeros(1, dtype=torch.int32, device=dev)
        self._lib = _lib
        self.hdl.barrier()  # signal pads are zero-initialised by torch; make sure every rank has mapped everything

    def __call__(self, flat: torch.Tensor):
        assert flat.dtype is torch.float32 and flat.is_contiguous() and flat.numel() <= self.cap
        L = self._lib.lib()
        self._lib.check(L.gode_allreduce_p2p(flat.data_ptr(), flat.numel(), self.hdl.buffer_ptrs_dev,
                                             self.hdl.signal_pad_ptrs_dev, self.rank, self.world, self.cap,
                                             self.epoch.data_ptr(), torch.cuda.current_stream().cuda_stream),
                        "gode_allreduce_p2p")
        return flat


class ExchangeMode(str):
    """What an enable_* call selected ('p2p', 'fused', 'nccl', 'off').  Truthy exactly when the requested peer-memory
    path is in place, so `if enable_…():` keeps working; compare or print it to see which path runs."""

    def __new__(cls, mode: str, ok: bool, why: str = ""):
        self = super().__new__(cls, mode)
        self.ok, self.why = ok, why
        return self

    def __bool__(self):
        return self.ok


def enable_p2p_allreduce(group=None, cap: int = 65536, require: bool = False) -> ExchangeMode:
    """Route the parameter-gradient exchange through the one-shot peer-memory kernel instead of NCCL.  Returns the mode
    now in place: 'p2p' (truthy), or 'nccl' (falsy; NCCL stays) if symmetric memory cannot be set up on this system —
    with require=True that case raises instead, so a caller that asked for peer memory never runs on NCCL unknowingly."""
    try:
        _api.config.grad_allreduce = P2PAllReduce(group, cap)
        return ExchangeMode("p2p", True)
    except Exception as e:  # noqa: BLE001
        if require:
            raise
        import sys
        sys.stderr.write("[gan_ode_b200] peer-memory all-reduce unavailable ({}); using NCCL\n".format(str(e)[:200]))
        _api.config.grad_allreduce = True if group is None else group
        return ExchangeMode("nccl", False, str(e)[:200])


class WorldNorm:
    """Exchange buffers for dopri5 with a world-scope error norm (gode_dopri5_fwd_world): every rank's step controller
    sees the RMS norm over the trajectories of all ranks, the partial sums travel as tagged words over NVLink peer memory
    inside the solver kernel.  torch's symmetric memory provides the peer mapping only."""

    def __init__(self, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        group = dist.group.WORLD if group is None else group
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 32:
            raise RuntimeError("world-scope norm supports up to 32 ranks")
        dev = torch.device("cuda", torch.cuda.current_device())
        words = 2 * 4 * self.world                      # include/gode.h GODE_WORLD_SLOT_WORDS
        self.buf = symm_mem.empty(2 * words, dtype=torch.int32, device=dev)   # uint64 words as pairs of int32
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, group)
        self.counter = torch.zeros(1, dtype=torch.int32, device=dev)
        self._totals = {}
        torch.cuda.synchronize()
        self.hdl.barrier()

    def total_batch(self, B: int) -> int:
        """Trajectories over all ranks when this rank holds B (one small all-reduce per distinct B, cached)."""
        v = self._totals.get(B)
        if v is None:
            t = torch.tensor([B], dtype=torch.int64, device=self.buf.device)
            dist.all_reduce(t, group=self.group)
            v = self._totals[B] = int(t.item())
        return v

    def struct(self, B: int):
        from . import _lib
        w = _lib.GodeWorld()
        w.rank, w.world, w.total_B = self.rank, self.world, self.total_batch(B)
        w.slots_dev, w.launch_ctr = self.hdl.buffer_ptrs_dev, self.counter.data_ptr()
        return w


def enable_world_norm(group=None) -> "WorldNorm":
    """Make options={'norm': 'world'} available to odeint / odeint_adjoint(method='dopri5') on this process group."""
    _api.config.world_norm = WorldNorm(group)
    return _api.config.world_norm


class FusedGradExchange:
    """Exchange buffers for the parameter-gradient all-reduce fused into the backward kernel's reduction tail
    (gode_dopri5_backprop_bwd_world): tagged words over NVLink peer memory, no separate all-reduce launch."""

    def __init__(self, group=None, max_params: int = 1088):
        import torch.distributed._symmetric_memory as symm_mem
        group = dist.group.WORLD if group is None else group
        self.world, self.rank, self.max_params = dist.get_world_size(group), dist.get_rank(group), int(max_params)
        dev = torch.device("cuda", torch.cuda.current_device())
        self._bufs = {}
        self._symm, self._group, self._dev = symm_mem, group, dev

    def struct(self, P: int):
        """GodeWorld for a gradient of P floats (one buffer + launch counter per distinct P, created collectively on first
        use: every rank reaches this call in the same order)."""
        from . import _lib
        ent = self._bufs.get(P)
        if ent is None:
            buf = self._symm.empty(2 * 2 * self.world * P, dtype=torch.int32, device=self._dev)   # uint64 words
            buf.zero_()
            hdl = self._symm.rendezvous(buf, self._group)
            ctr = torch.zeros(1, dtype=torch.int32, device=self._dev)
            torch.cuda.synchronize()
            hdl.barrier()
            ent = self._bufs[P] = (buf, hdl, ctr)
        w = _lib.GodeWorld()
        w.rank, w.world, w.total_B = self.rank, self.world, 0
        w.slots_dev, w.launch_ctr = ent[1].buffer_ptrs_dev, ent[2].data_ptr()
        return w


def enable_fused_grad_exchange(group=None, require: bool = False) -> ExchangeMode:
    """dopri5 backprop-through-solver: all-reduce the parameter gradient inside the backward kernel.  Other backward kernels
    keep using config.grad_allreduce (set it as well).  Returns 'fused' (truthy) or 'off' (falsy: symmetric memory is
    unavailable, config.grad_exchange stays None); require=True raises in the second case."""
    try:
        ex = FusedGradExchange(group)
        ex.struct(544)      # the reference shape: create its buffers now, outside any graph capture
        _api.config.grad_exchange = ex
        return ExchangeMode("fused", True)
    except Exception as e:  # noqa: BLE001
        if require:
            raise
        import sys
        sys.stderr.write("[gan_ode_b200] fused gradient exchange unavailable ({})\n".format(str(e)[:200]))
        _api.config.grad_exchange = None
        return ExchangeMode("off", False, str(e)[:200])
