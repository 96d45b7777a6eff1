"""CPU restatement of torchsde's fixed-step Euler–Maruyama driver for the SDE sampler.

TEST INFRASTRUCTURE ONLY (see oracle/torchdiffeq_restatement.py header; same rules).
PARITY UNPINNED: torchsde (requirements.txt:13, unpinned) is not vendored or installable
here; this restates its published algorithm:
  torchsde/_core/sdeint.py::sdeint          front door, ts handling
  torchsde/_core/base_solver.py::integrate  fixed-step loop with fp32 time accumulation and
                                            linear interpolation onto the output times
  torchsde/_core/methods/euler.py::step     y1 = y0 + f dt + g (.) dW   (diagonal noise, Ito)
  torchsde/_core/interp.py::linear_interp
Reference call site: models/mocogan_sde.py:57-59
  sdeint_adjoint(SDEFunc, x, linspace(0,1,T), method='euler', adjoint_method='euler', dt=2.5e-2)

torchsde's BrownianInterval draws are not reproducible from a counter stream and the reference
passes no `bm`/seed, so parity is pinned GIVEN the increments: `bm` here is any
callable `bm(t0, t1) -> (B, D)` tensor.  `step_grid(ts, dt)` exposes the exact (t0, t1) pairs
the driver visits, so a test (or the CUDA path's Philox stream) can supply one increment per step.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch

__all__ = ["sdeint", "step_grid", "TableBrownian"]


def _linear_interp(t0, y0, t1, y1, t):
    """interp.py::linear_interp."""
    assert t0 <= t <= t1, "Incorrect time order for linear interpolation: t0={}, t={}, t1={}.".format(t0, t, t1)
    y = (t1 - t) / (t1 - t0) * y0 + (t - t0) / (t1 - t0) * y1
    return y


def step_grid(ts: torch.Tensor, dt: float) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """The (curr_t, next_t) pairs base_solver.integrate visits — 0-d tensors in ts.dtype, so the
    accumulation `curr_t + dt` rounds in fp32 exactly as upstream (41 steps for linspace(0,1,16),
    dt=0.025; the last one is 4.2e-7 long)."""
    pairs = []
    curr_t = ts[0]
    for out_t in ts[1:]:
        while curr_t < out_t:
            next_t = min(curr_t + dt, ts[-1])
            pairs.append((curr_t, next_t))
            curr_t = next_t
    return pairs


class TableBrownian:
    """A `bm` that replays a table of per-step increments in visiting order."""

    def __init__(self, increments: torch.Tensor):
        self.increments = increments  # (n_steps, B, D)
        self.k = 0

    def __call__(self, t0, t1):
        w = self.increments[self.k]
        self.k += 1
        return w


def sdeint(sde, y0, ts, bm: Optional[Callable] = None, method=None, dt=1e-3, adaptive=False, **unused):
    """sdeint.py::sdeint for method='euler', noise_type='diagonal', sde_type='ito', adaptive=False.
    Differentiable by ordinary autograd (discretise-then-optimise given the increments)."""
    if method not in (None, "euler"):
        raise ValueError("oracle restates method='euler' only")
    if adaptive:
        raise NotImplementedError("adaptive SDE stepping is not on the gan-ode hot path")
    assert getattr(sde, "noise_type", "diagonal") == "diagonal" and getattr(sde, "sde_type", "ito") == "ito"
    if not torch.is_tensor(ts):
        ts = torch.tensor(ts, dtype=y0.dtype, device=y0.device)
    assert (ts[1:] > ts[:-1]).all(), "ts must be strictly increasing"
    if bm is None:
        gen = torch.Generator().manual_seed(0)

        def bm(t0, t1):
            return torch.randn(y0.shape, generator=gen, dtype=y0.dtype) * float(t1 - t0) ** 0.5

    step_size = dt
    prev_t = curr_t = ts[0]
    prev_y = curr_y = y0
    ys = [y0]
    for out_t in ts[1:]:
        while curr_t < out_t:
            next_t = min(curr_t + step_size, ts[-1])
            prev_t, prev_y = curr_t, curr_y
            # euler.py::Euler.step
            h = next_t - curr_t
            I_k = bm(curr_t, next_t)
            f = sde.f(curr_t, curr_y)
            g_prod = sde.g(curr_t, curr_y) * I_k
            curr_y = curr_y + f * h + g_prod
            curr_t = next_t
        ys.append(_linear_interp(t0=prev_t, y0=prev_y, t1=curr_t, y1=curr_y, t=out_t))
    return torch.stack(ys, dim=0)
