"""torchsde-signature boundary over the fused Euler–Maruyama kernels.

Mirrors what the reference imports (`from torchsde import sdeint_adjoint as sdeint`, models/mocogan_sde.py:4) for the
one way it is called (models/mocogan_sde.py:57-59):

    sdeint(self.ode_fn, x, torch.linspace(0, 1, T).float(), method='euler', adjoint_method='euler', dt=2.5e-2)

`sde` must be the reference's SDEFunc structure (models/mocogan_sde.py:6-27): `drift_fn` and `diffusion_fn` are
Sequential(Linear(D,H), Tanh(), Linear(H,D)), noise_type 'diagonal', sde_type 'ito'.  The whole solve (every step of
every trajectory, both MLPs, the Brownian increments, the frame interpolation) is one kernel launch; there is no
eager fallback.

Brownian motion (`bm`):
  None                 -> PhiloxBrownian with a seed drawn from torch's CPU generator (reproducible under manual_seed)
  PhiloxBrownian(seed) -> counter-based Philox4x32-10 increments generated inside the kernels, regenerated (not stored)
                          by the backward, keyed by GLOBAL trajectory index so results do not depend on sharding
  TableBrownian(dW)    -> a given (n_steps, B, D) table of increments (parity with a prescribed path)
torchsde's own BrownianInterval objects are not accepted (their draws are not reproducible from a counter).

Backward: `sdeint` and `sdeint_adjoint` both return the exact discrete gradient of the Euler–Maruyama recursion for
the same increments (backprop-through-solver).  torchsde's stochastic adjoint re-solve is SURVEY §8f-3.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import GodeError
import importlib

_api = importlib.import_module(__package__ + ".odeint")

__all__ = ["sdeint", "sdeint_adjoint", "PhiloxBrownian", "TableBrownian", "step_grid", "recognise_sde"]


class PhiloxBrownian:
    def __init__(self, seed: int, traj_offset: int = 0):
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.traj_offset = int(traj_offset)


class TableBrownian:
    def __init__(self, increments: torch.Tensor):
        self.increments = increments


def _mlp(seq):
    ok = (isinstance(seq, nn.Sequential) and len(seq) == 3 and isinstance(seq[0], nn.Linear)
          and isinstance(seq[1], nn.Tanh) and isinstance(seq[2], nn.Linear)
          and seq[0].bias is not None and seq[2].bias is not None)
    return (seq[0].weight, seq[0].bias, seq[2].weight, seq[2].bias) if ok else None


def recognise_sde(sde):
    f, g = _mlp(getattr(sde, "drift_fn", None)), _mlp(getattr(sde, "diffusion_fn", None))
    if f is None or g is None or f[0].shape != g[0].shape or f[2].shape != g[2].shape:
        raise NotImplementedError("gan_ode_b200 fuses the solver with the reference's SDEFunc (models/mocogan_sde.py:6-27): "
                                  "sde.drift_fn and sde.diffusion_fn must be Sequential(Linear(D,H), Tanh(), Linear(H,D))")
    if getattr(sde, "noise_type", "diagonal") != "diagonal" or getattr(sde, "sde_type", "ito") != "ito":
        raise NotImplementedError("only noise_type='diagonal', sde_type='ito' (the reference's) are supported")
    return f, g


def step_grid(ts: torch.Tensor, dt: float):
    """The fixed-step grid of torchsde's base_solver.integrate in fp32 (curr_t is a 0-d fp32 tensor upstream):
    returns (h[n_steps], out_step[T], w0[T], w1[T]) — frame j = w0[j]*y_k + w1[j]*y_{k+1} with k = out_step[j]."""
    t = ts.detach().cpu().numpy().astype(np.float32)
    assert (np.diff(t) > 0).all(), "ts must be strictly increasing"
    T = len(t)
    dt32 = np.float32(dt)
    h, out_step = [], np.zeros(T, dtype=np.int32)
    w0, w1 = np.zeros(T, dtype=np.float32), np.zeros(T, dtype=np.float32)
    curr = prev = t[0]
    for j in range(1, T):
        while curr < t[j]:
            nxt = min(np.float32(curr + dt32), t[-1])
            prev, curr = curr, nxt
            h.append(np.float32(curr - prev))
        assert len(h) > 0 and prev <= t[j] <= curr
        out_step[j] = len(h) - 1
        w0[j] = np.float32(curr - t[j]) / np.float32(curr - prev)
        w1[j] = np.float32(t[j] - prev) / np.float32(curr - prev)
    return np.asarray(h, dtype=np.float32), out_step, w0, w1


def _ptrs(ws):
    return (C.c_void_p * 4)(*[w.data_ptr() for w in ws])


class _SdeEM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y0, meta, *weights):
        L = _lib.lib()
        B, D = y0.shape
        H = weights[0].shape[0]
        h, out_step, w0, w1 = meta["grid"]
        T, n_steps = len(out_step), len(h)
        ws = [_api._f32c(w) for w in weights]
        y0c = _api._f32c(y0)
        buf, view = _api._alloc_traj(T, B, D, meta["layout"], y0)
        keep = meta["keep"]  # decided outside: grad mode is off inside Function.forward
        states = torch.empty((n_steps, B, D), dtype=torch.float32, device=y0.device) if keep else None
        dW = meta["dW"]
        rc = L.gode_sde_em_fwd(y0c.data_ptr(), _ptrs(ws[:4]), _ptrs(ws[4:]), h.ctypes.data, n_steps, out_step.ctypes.data,
                               w0.ctypes.data, w1.ctypes.data, B, D, H, T, None if dW is None else dW.data_ptr(),
                               meta["seed"], meta["traj_offset"], meta["layout"], buf.data_ptr(),
                               None if states is None else states.data_ptr(), _api._stream())
        _lib.check(rc, "gode_sde_em_fwd")
        ctx.meta = meta
        ctx.save_for_backward(states, *ws)
        return view

    @staticmethod
    @_api._bwd_on_device
    def backward(ctx, grad_frames):
        L = _lib.lib()
        states, *ws = ctx.saved_tensors
        if states is None:
            raise GodeError("SDE forward ran without keeping states (inputs did not require grad)")
        meta = ctx.meta
        h, out_step, w0, w1 = meta["grid"]
        n_steps, B, D = states.shape
        H, T = ws[0].shape[0], len(out_step)
        g = _api._grad_in_layout(grad_frames, meta["layout"])
        P = L.gode_param_count(D, H)
        grad_y0 = torch.empty((B, D), dtype=torch.float32, device=states.device)
        grad_p = torch.empty(2 * P, dtype=torch.float32, device=states.device)
        ws_bytes = L.gode_sde_workspace_bytes(B, D, H)
        wsp = _api._workspace(states.device, ws_bytes)
        dW = meta["dW"]
        rc = L.gode_sde_em_bwd(states.data_ptr(), g.data_ptr(), _ptrs(ws[:4]), _ptrs(ws[4:]), h.ctypes.data, n_steps,
                               out_step.ctypes.data, w0.ctypes.data, w1.ctypes.data, B, D, H, T,
                               None if dW is None else dW.data_ptr(), meta["seed"], meta["traj_offset"], meta["layout"],
                               grad_y0.data_ptr(), grad_p.data_ptr(), wsp.data_ptr(), ws_bytes, _api._stream())
        _lib.check(rc, "gode_sde_em_bwd")
        _api._maybe_allreduce(grad_p)
        needs = ctx.needs_input_grad
        n1 = H * D
        outs = []
        for base in (0, P):
            outs += [grad_p[base:base + n1].view(H, D), grad_p[base + n1:base + n1 + H],
                     grad_p[base + n1 + H:base + n1 + H + D * H].view(D, H), grad_p[base + n1 + H + D * H:base + P]]
        outs = [o if need else None for o, need in zip(outs, needs[2:10])]
        return ((grad_y0 if needs[0] else None), None, *outs)


def _solve(sde, y0, ts, bm, method, dt, adaptive, options, force_keep=False):
    f, g = recognise_sde(sde)
    if method not in (None, "euler"):
        raise NotImplementedError("only method='euler' (the reference's, models/mocogan_sde.py:58) is on the hot path")
    if adaptive:
        raise NotImplementedError("adaptive SDE stepping is not on the gan-ode hot path")
    if not isinstance(y0, torch.Tensor) or not torch.is_floating_point(y0):
        raise TypeError("`y0` must be a floating point Tensor")
    if y0.dim() != 2 or y0.dtype != torch.float32:
        raise NotImplementedError("y0 must be a (B, D) float32 tensor")
    _api._require_cuda(y0, weights=(*f, *g))
    if not torch.is_tensor(ts):
        ts = torch.tensor(ts, dtype=y0.dtype)
    D, H = f[0].shape[1], f[0].shape[0]
    if not _lib.lib().gode_supported(D, H, _lib.PREC["fp32"]):
        raise NotImplementedError("no sm_100a kernel compiled for SDEFunc(dim={}, dim_hidden={})".format(D, H))
    options = {} if options is None else dict(options)
    grid = step_grid(ts, dt)
    keep = force_keep or (torch.is_grad_enabled() and (y0.requires_grad or any(w.requires_grad for w in (*f, *g))))
    meta = dict(grid=grid, layout=_api._layout_code(options.get("layout", _api.config.layout)), dW=None, seed=0,
                traj_offset=0, keep=keep)
    if bm is None:
        bm = PhiloxBrownian(int(torch.randint(0, 2 ** 62, (1,)).item()))
    if isinstance(bm, PhiloxBrownian):
        meta["seed"], meta["traj_offset"] = bm.seed, bm.traj_offset
    elif isinstance(bm, TableBrownian):
        dW = _api._f32c(bm.increments.to(y0.device))
        if tuple(dW.shape) != (len(grid[0]), y0.shape[0], D):
            raise ValueError("increments must have shape (n_steps={}, B, D); got {}".format(len(grid[0]), tuple(dW.shape)))
        meta["dW"] = dW
    else:
        raise NotImplementedError("bm must be None, PhiloxBrownian or TableBrownian (torchsde BrownianInterval objects "
                                  "are not reproducible from a counter stream)")
    with _api._on_device(y0.device):
        return _SdeEM.apply(y0, meta, *f, *g)


def sdeint(sde, y0, ts, bm=None, method=None, dt=1e-3, adaptive=False, rtol=1e-5, atol=1e-4, dt_min=1e-5, options=None,
           names=None, logqp=False, extra=False, extra_solver_state=None, **unused_kwargs):
    """torchsde.sdeint for the reference's SDEFunc (Euler–Maruyama, diagonal Ito noise)."""
    if names or logqp or extra or extra_solver_state is not None:
        raise NotImplementedError("names / logqp / extra are not on the gan-ode hot path")
    return _solve(sde, y0, ts, bm, method, dt, adaptive, options)


def sdeint_adjoint(sde, y0, ts, bm=None, method=None, adjoint_method=None, dt=1e-3, adaptive=False,
                   adjoint_adaptive=False, rtol=1e-5, adjoint_rtol=1e-5, atol=1e-4, adjoint_atol=1e-4, dt_min=1e-5,
                   options=None, adjoint_options=None, adjoint_params=None, names=None, logqp=False, extra=False,
                   extra_solver_state=None, **unused_kwargs):
    """torchsde.sdeint_adjoint signature (models/mocogan_sde.py:57-59).  Gradient = exact discrete backprop through
    the Euler–Maruyama steps with regenerated Philox increments (see module docstring)."""
    if names or logqp or extra or extra_solver_state is not None:
        raise NotImplementedError("names / logqp / extra are not on the gan-ode hot path")
    if adjoint_method not in (None, "euler") or adjoint_adaptive:
        raise NotImplementedError("adjoint_method must be 'euler' (the reference's)")
    return _solve(sde, y0, ts, bm, method, dt, adaptive, options)
