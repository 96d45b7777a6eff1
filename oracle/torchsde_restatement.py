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


# =====================================================================================================================
# torchsde's stochastic adjoint (reference call: models/mocogan_sde.py:57-59, adjoint_method='euler')
#   torchsde/_core/adjoint.py::_SdeintAdjointMethod          forward without a graph; backward = per output interval,
#                                                            last to first, ONE solve of the augmented adjoint SDE on
#                                                            [-t_i, -t_{i-1}] with the reversed Brownian motion and the same
#                                                            dt, then y <- ys[i-1], a <- a + grad_ys[i-1]
#   torchsde/_core/adjoint_sde.py::AdjointSDE                f_and_g_prod for (ito, diagonal) = f_corrected_diagonal + g_prod
#   torchsde/_brownian/derived.py::ReverseBrownian           bm_rev(ta, tb) = bm(-tb, -ta)
# [TS-recalled]: restated from the published source (torchsde 0.2.x); PARITY UNPINNED like the rest of this file.
#
# The adjoint queries the Brownian motion on intervals that are NOT forward steps (the reverse grid restarts at every frame
# time), so `bm` must answer arbitrary (ta, tb) consistently with the forward increments.  `GridBrownian` is a Brownian path
# sampled on the UNION of the forward and reverse step times (`adjoint_time_grid`): an increment over any interval whose end
# points are grid times is the left-to-right fp32 sum of the grid increments inside it.
# =====================================================================================================================
def reverse_step_grid(ts: torch.Tensor, dt: float):
    """Per output interval i = T-1..1 the (s0, s1) pairs base_solver.integrate visits on [-ts[i], -ts[i-1]] (0-d tensors in
    ts.dtype, fp32 accumulation as upstream): 3 steps of 0.025, 0.025, 0.01667 per interval for the reference call, 45 in all."""
    out = []
    for i in range(len(ts) - 1, 0, -1):
        pairs = []
        curr, end = -ts[i], -ts[i - 1]
        while curr < end:
            nxt = min(curr + dt, end)
            pairs.append((curr, nxt))
            curr = nxt
        out.append((i, pairs))
    return out


def adjoint_time_grid(ts: torch.Tensor, dt: float) -> torch.Tensor:
    """Sorted unique union (float64 copies of the fp32 values) of every time the forward solve and the adjoint solves visit."""
    pts = {float(ts[0])}
    for a, b in step_grid(ts, dt):
        pts.add(float(a)); pts.add(float(b))
    for _, pairs in reverse_step_grid(ts, dt):
        for s0, s1 in pairs:
            pts.add(float(-s0)); pts.add(float(-s1))
    return torch.tensor(sorted(pts), dtype=torch.float64)


class GridBrownian:
    """W sampled on `times` (R+1 points): `increments[r] = W(times[r+1]) - W(times[r])`, shape (R, B, D).
    bm(ta, tb) = W(tb) - W(ta) for grid times ta < tb: the fp32 left-to-right sum of the increments in between."""

    def __init__(self, times: torch.Tensor, increments: torch.Tensor):
        assert increments.shape[0] == len(times) - 1
        self.times, self.increments = times.double(), increments

    def index(self, t) -> int:
        k = int(torch.argmin((self.times - float(t)).abs()))
        assert abs(float(self.times[k]) - float(t)) <= 1e-9, "time {} is not on the Brownian grid".format(float(t))
        return k

    def __call__(self, ta, tb):
        ia, ib = self.index(ta), self.index(tb)
        assert ia < ib
        w = self.increments[ia]
        for r in range(ia + 1, ib):
            w = w + self.increments[r]
        return w


def _vjp(outputs, inputs, grad_outputs, **kw):
    """torchsde/_core/misc.py::vjp: autograd.grad with allow_unused, None -> zeros."""
    gs = torch.autograd.grad(outputs, inputs, grad_outputs, allow_unused=True, **kw)
    return [torch.zeros_like(x) if g is None else g for g, x in zip(gs, inputs)]


def adjoint_f_and_g_prod(sde, params, t, y, adj_y, v):
    """adjoint_sde.py::AdjointSDE.f_and_g_prod_corrected_diagonal at reverse time t (forward time -t), on detached state.
    Returns (f_out, g_out): tuples (dy, da, dtheta...) — the drift and the diffusion-times-increment of the augmented state."""
    y = y.detach().requires_grad_(True)
    adj_y = adj_y.detach()
    with torch.enable_grad():
        f = sde.f(-t, y)
        g = sde.g(-t, y)
        g_prod = g * v
        # _f_corrected_diagonal
        g_dg_vjp, = _vjp(g, [y], g, create_graph=True)
        f_corr = f - g_dg_vjp                       # "double Stratonovich correction"
        vjp_y_and_params = _vjp(f_corr, [y] + params, adj_y, retain_graph=True)
        a_dg_vjp, = _vjp(g, [y], adj_y, retain_graph=True)          # back to Ito form
        extra = _vjp(g, [y] + params, a_dg_vjp.detach(), retain_graph=True)
        f_out = [-f_corr.detach()] + [a + b for a, b in zip(vjp_y_and_params, extra)]
        # _g_prod
        g_out = [-g_prod.detach()] + _vjp(g_prod, [y] + params, adj_y)
    return f_out, g_out


class _SdeintAdjoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sde, ts, dt, bm, n_params, y0, *params):
        with torch.no_grad():
            ys = sdeint(sde, y0.detach(), ts, bm=bm, method="euler", dt=dt)
        ctx.sde, ctx.dt, ctx.bm, ctx.ts = sde, dt, bm, ts
        ctx.save_for_backward(ys, *params)
        return ys

    @staticmethod
    def backward(ctx, grad_ys):
        ys, *params = ctx.saved_tensors
        sde, dt, bm, ts = ctx.sde, ctx.dt, ctx.bm, ctx.ts
        params = list(params)
        y, a = ys[-1], grad_ys[-1]
        theta = [torch.zeros_like(p) for p in params]
        with torch.no_grad():
            for i in range(len(ts) - 1, 0, -1):
                curr, end = -ts[i], -ts[i - 1]
                while curr < end:                                   # base_solver.integrate, fixed step
                    nxt = min(curr + dt, end)
                    h = nxt - curr
                    v = bm(-nxt, -curr)                             # ReverseBrownian(bm)(curr, nxt)
                    f_out, g_out = adjoint_f_and_g_prod(sde, params, curr, y, a, v)
                    # euler.py::Euler.step on the flattened augmented state: y1 = y0 + f * dt + g_prod
                    y = y + f_out[0] * h + g_out[0]
                    a = a + f_out[1] * h + g_out[1]
                    theta = [th + fo * h + go for th, fo, go in zip(theta, f_out[2:], g_out[2:])]
                    curr = nxt
                y = ys[i - 1]
                a = a + grad_ys[i - 1]
        return (None, None, None, None, None, a, *theta)


def sdeint_adjoint(sde, y0, ts, bm, method=None, adjoint_method=None, dt=1e-3, adaptive=False, **unused):
    """adjoint.py::sdeint_adjoint for method = adjoint_method = 'euler', (ito, diagonal), fixed step.  `bm` must answer
    arbitrary grid-time queries (GridBrownian)."""
    if method not in (None, "euler") or adjoint_method not in (None, "euler") or adaptive:
        raise NotImplementedError("oracle restates the reference's call only (euler / euler, fixed step)")
    if not torch.is_tensor(ts):
        ts = torch.tensor(ts, dtype=y0.dtype)
    params = [p for p in sde.parameters() if p.requires_grad]
    return _SdeintAdjoint.apply(sde, ts, dt, bm, len(params), y0, *params)
