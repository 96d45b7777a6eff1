"""ODE-RNN latent-motion sampler (models/mocogan_ode_rnn.py:40-54) on the fused kernels, one C-ABI call per direction.

The reference's own loop (`for frame: h' = odeint(ode_fn, h, [0,1])[-1]; h = GRUCell(e_t, h')`) already runs unchanged
through `install_shims()` — 16 fused dopri5 launches plus `nn.GRUCell` in PyTorch.  This module is the opt-in fused path
for the same computation: `odernn_codes(ode_fn, gru, h0, eps)` enqueues all F (solve, jump) pairs from C
(`gode_odernn_fwd`), the GRU jump is a CUDA kernel of this library, and the backward (`gode_odernn_bwd`) walks the frames
in reverse on the device.  Step control, tolerances and results are those of F separate `odeint` calls with torchdiffeq's
defaults; gradients are the discrete adjoint of the recorded solves (see odeint._solve), or with
`options={'adjoint': 'continuous'}` torchdiffeq's continuous adjoint per frame, as the reference loop computes them.  `options={'norm': 'trajectory'}`
(opt-in, not what torchdiffeq does for a batch) gives every trajectory its own step control: no grid-wide reduction per
attempted step, ordinary launches, any batch size.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from ._lib import GodeAdaptiveOpts
import importlib

_api = importlib.import_module(__package__ + ".odeint")   # (the package re-exports the FUNCTION odeint under that name)
from .odeint import _adaptive_opts, _f32c, _ptr, _stream, config, recognise_field

__all__ = ["odernn_codes", "gru_jump", "OdeRnnLog"]


def _gru_params(cell):
    if not isinstance(cell, nn.GRUCell) or not cell.bias:
        raise NotImplementedError("the fused jump is nn.GRUCell with bias (models/mocogan.py:198); got {!r}".format(cell))
    return cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh


class _GruJump(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, h, w_ih, w_hh, b_ih, b_hh):
        L = _lib.lib()
        B, D = h.shape
        xs = [_f32c(v) for v in (x, h, w_ih, w_hh, b_ih, b_hh)]
        out = torch.empty_like(xs[1])
        _lib.check(L.gode_gru_jump_fwd(*[v.data_ptr() for v in xs], B, D, out.data_ptr(), _stream()), "gode_gru_jump_fwd")
        ctx.save_for_backward(*xs)
        return out

    @staticmethod
    @_api._bwd_on_device
    def backward(ctx, go):
        L = _lib.lib()
        xs = ctx.saved_tensors
        B, D = xs[1].shape
        go = _f32c(go)
        gx, gh = torch.empty_like(xs[0]), torch.empty_like(xs[1])
        gp = torch.empty(L.gode_gru_param_count(D), dtype=torch.float32, device=go.device)
        wsb = L.gode_odernn_workspace_bytes(B, D, D)
        ws = torch.empty(wsb, dtype=torch.uint8, device=go.device)
        _lib.check(L.gode_gru_jump_bwd(*[v.data_ptr() for v in xs], go.data_ptr(), B, D, gx.data_ptr(), gh.data_ptr(),
                                       gp.data_ptr(), ws.data_ptr(), wsb, _stream()), "gode_gru_jump_bwd")
        n = 3 * D * D
        return gx, gh, gp[:n].view(3 * D, D), gp[n:2 * n].view(3 * D, D), gp[2 * n:2 * n + 3 * D], gp[2 * n + 3 * D:]


def gru_jump(x, h, cell):
    """`cell(x, h)` for an nn.GRUCell (the ODE-RNN jump, models/mocogan_ode_rnn.py:49) as one fused kernel each way."""
    gp = _gru_params(cell)
    _api._require_cuda(x, h, what="x / h", weights=gp)
    with _api._on_device(h.device):
        return _GruJump.apply(x, h, *gp)


class OdeRnnLog:
    """Host view of the per-frame step logs of the last fused sampler call (reading synchronises)."""

    def __init__(self, raw, stride, F):
        self._raw, self._stride, self._F = raw, stride, F

    def frames(self):
        h = self._raw.cpu().numpy().tobytes()
        out = []
        for f in range(self._F):
            hdr = _lib.GodeStepLog.from_buffer_copy(h[f * self._stride:f * self._stride + C.sizeof(_lib.GodeStepLog)])
            out.append(dict(status=hdr.status, n_attempts=hdr.n_attempts, n_accepted=hdr.n_accepted, nfe=hdr.nfe, dt0=hdr.dt0))
        return out


_LAST = [None]


def last_log():
    return _LAST[0]


class _OdeRnn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h0, eps, meta, W1, b1, W2, b2, w_ih, w_hh, b_ih, b_hh):
        L = _lib.lib()
        F, B, D = eps.shape
        H = W1.shape[0]
        dev = h0.device
        ts = [_f32c(v) for v in (h0, eps, W1, b1, W2, b2, w_ih, w_hh, b_ih, b_hh)]
        o = meta["opts"]
        keep = meta["keep"] and meta["adj_opts"] is None    # the continuous adjoint needs the frame end points only
        opts = GodeAdaptiveOpts.from_buffer_copy(bytes(o))
        kc = o.ckpt_capacity if keep else 0
        opts.ckpt_capacity = kc
        stride = L.gode_odernn_log_stride(o.log_capacity)
        codes = torch.empty((F, B, D), dtype=torch.float32, device=dev)
        seg = torch.empty((F, 2, B, D), dtype=torch.float32, device=dev)
        logs = torch.empty(F * stride, dtype=torch.uint8, device=dev)
        per_traj = o.norm_scope == _lib.NORM_TRAJ
        ckpt = torch.empty((F, max(kc, 1), B, D), dtype=torch.float32, device=dev) if keep else None
        acc = (torch.empty((F, 2, max(kc, 1)) + ((B,) if per_traj else ()), dtype=torch.float64, device=dev)) if keep else None
        n_acc = torch.empty((F + 1, B), dtype=torch.int32, device=dev) if per_traj else None
        wsb = L.gode_odernn_workspace_bytes(B, D, H)
        ws = _api._workspace(dev, wsb)
        _lib.check(L.gode_odernn_fwd(*[v.data_ptr() for v in ts], B, D, H, F, C.byref(opts), codes.data_ptr(),
                                     seg.data_ptr(), logs.data_ptr(), _ptr(ckpt), _ptr(acc), _ptr(n_acc), ws.data_ptr(), wsb,
                                     _stream()),
                   "gode_odernn_fwd")
        _LAST[0] = OdeRnnLog(logs, stride, F)
        ctx.meta, ctx.kc = meta, kc
        ctx.save_for_backward(seg, logs, ckpt, acc, n_acc, *ts[1:])
        return codes

    @staticmethod
    @_api._bwd_on_device
    def backward(ctx, grad_codes):
        L = _lib.lib()
        seg, logs, ckpt, acc, n_acc, eps, W1, b1, W2, b2, w_ih, w_hh, b_ih, b_hh = ctx.saved_tensors
        ao = ctx.meta["adj_opts"]
        if ckpt is None and ao is None:
            raise _lib.GodeError("ODE-RNN forward ran without checkpoints (inputs did not require grad)")
        F, _, B, D = seg.shape
        H = W1.shape[0]
        dev = seg.device
        g = _f32c(grad_codes)
        P1, P2 = L.gode_param_count(D, H), L.gode_gru_param_count(D)
        gh0 = torch.empty((B, D), dtype=torch.float32, device=dev)
        need_eps = ctx.needs_input_grad[1]
        geps = torch.empty((F, B, D), dtype=torch.float32, device=dev) if need_eps else None
        gode_, ggru = torch.empty(P1, dtype=torch.float32, device=dev), torch.empty(P2, dtype=torch.float32, device=dev)
        scratch = torch.empty(3 * B * D + F * P1, dtype=torch.float32, device=dev)
        wsb = L.gode_odernn_workspace_bytes(B, D, H)
        ws = _api._workspace(dev, wsb)
        o = ctx.meta["opts"]
        _lib.check(L.gode_odernn_bwd(g.data_ptr(), eps.data_ptr(), W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(),
                                     w_ih.data_ptr(), w_hh.data_ptr(), b_ih.data_ptr(), b_hh.data_ptr(), B, D, H, F,
                                     o.log_capacity, ctx.kc, seg.data_ptr(), logs.data_ptr(), _ptr(ckpt), _ptr(acc),
                                     _ptr(n_acc), C.byref(ao) if ao is not None else None, ctx.meta["param_mask"], gh0.data_ptr(), _ptr(geps), gode_.data_ptr(), ggru.data_ptr(), scratch.data_ptr(),
                                     ws.data_ptr(), wsb, _stream()), "gode_odernn_bwd")
        from .odeint import _maybe_allreduce
        _maybe_allreduce(gode_)
        _maybe_allreduce(ggru)
        n1, n = H * D, 3 * D * D
        return (gh0, geps, None,
                gode_[:n1].view(H, D), gode_[n1:n1 + H], gode_[n1 + H:n1 + H + D * H].view(D, H), gode_[n1 + H + D * H:],
                ggru[:n].view(3 * D, D), ggru[n:2 * n].view(3 * D, D), ggru[2 * n:2 * n + 3 * D], ggru[2 * n + 3 * D:])


def odernn_codes(ode_fn, gru_cell, h0, eps, *, rtol=1e-7, atol=1e-9, options=None):
    """All frames of models/mocogan_ode_rnn.py:45-50 in one call.  h0: (B, D) initial hidden state, eps: (F, B, D) the
    per-frame noise inputs e_t (models/mocogan.py:297-301).  Returns the hidden states (F, B, D);
    `codes.transpose(0, 1).reshape(-1, D)` is the reference's `torch.cat(z_m_t[1:], dim=1).view(-1, D)` (:51-52)."""
    W1, b1, W2, b2 = recognise_field(ode_fn)
    gp = _gru_params(gru_cell)
    _api._require_cuda(h0, eps, what="h0 / eps", weights=(W1, b1, W2, b2) + tuple(gp))
    if h0.dim() != 2 or eps.dim() != 3 or eps.shape[1:] != h0.shape:
        raise ValueError("h0 must be (B, D) and eps (F, B, D)")
    D, H = W1.shape[1], W1.shape[0]
    if (D, H) != (16, 16) or gru_cell.input_size != D or gru_cell.hidden_size != D:
        raise NotImplementedError("the fused ODE-RNN kernels exist for the reference shape D=H=16, GRUCell(16,16)")
    options = {} if options is None else dict(options)
    o = _adaptive_opts(rtol, atol, options, 1.0)
    keep = torch.is_grad_enabled() and any(t.requires_grad for t in (h0, eps, W1, b1, W2, b2) + tuple(gp))
    mode = options.get("adjoint", "discrete")
    if mode not in ("continuous", "discrete"):
        raise ValueError("options['adjoint'] must be 'continuous' or 'discrete'")
    adj = None
    if mode == "continuous":
        if o.norm_scope != _lib.NORM_BATCH:
            raise NotImplementedError("the continuous adjoint uses torchdiffeq's batch-global norm")
        adj = _adaptive_opts(rtol, atol, {k: v for k, v in options.items() if k != "norm"}, 1.0)
    # adjoint.py: adjoint_params = the parameters with requires_grad; only those are in the augmented state and its norm (the
    # single-layer field's W2 = I, b2 = 0 are constants, frozen parameters drop out) — as odeint_adjoint derives it
    mask = sum(1 << k for k, q in enumerate((W1, b1, W2, b2)) if q.requires_grad)
    meta = dict(opts=o, keep=keep, adj_opts=adj, param_mask=mask)
    with _api._on_device(h0.device):
        return _OdeRnn.apply(h0, eps, meta, W1, b1, W2, b2, *gp)
