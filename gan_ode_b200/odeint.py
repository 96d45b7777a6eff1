"""torchdiffeq-signature boundary over the fused sm_100a kernels.

Mirrors the interface the reference imports (`from torchdiffeq import odeint_adjoint as odeint`,
models/mocogan_ode.py:4, models/mocogan_ode_rnn.py:4) so `sample_z_m` (models/mocogan_ode.py:133-148) runs
unchanged: same argument names, defaults (method=None -> dopri5, rtol=1e-7, atol=1e-9), error behaviour
(TypeError for non-floating y0 / t, AssertionError texts of torchdiffeq's solver asserts) and return value
((len(t), *y0.shape), sol[0] == y0 bit-exact, gradients to y0 and to func.parameters()).

What is NOT generic: `func` must be the reference's ODEFunc structure (models/mocogan_ode.py:6-17) —
`func.fn == Sequential(Linear(D,H), Tanh(), Linear(H,D))`, autonomous — because the whole solve (all steps, all
stages, the MLP inside each stage) is ONE kernel launch.  Anything else raises NotImplementedError; there is no
eager/CPU fallback.

Extra keys accepted in `options` (ignored by torchdiffeq, so reference code never passes them):
  precision : 'fp32' (default) | 'tf32' | 'bf16'      arithmetic of the MLP contractions
  norm      : 'batch' (default) | 'trajectory'        dopri5 step control: torchdiffeq's one-(t,dt)-per-batch, or
                                                      per-trajectory (= torchdiffeq called with B=1 per trajectory)
  layout    : 'tbd' (default) | 'btd'                 memory order of the returned (T,B,D) tensor; 'btd' makes the
                                                      caller's .transpose(0,1).reshape(-1,D) a free view
  check     : bool (default False)                    synchronise and raise solver asserts eagerly
  bwd_precision : 'bf16' | 'fp32'                     bf16 mode + odeint_adjoint: adjoint on tcgen05 or on the FP32 kernels
                                                      (default: 'bf16' for D=64/H=256, 'fp32' for D=H=16)
"""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from ._lib import GodeAdaptiveOpts, GodeError, GodeStepLog

__all__ = ["odeint", "odeint_adjoint", "last_step_log", "StepLog", "recognise_field", "config", "check_status"]


class _Config:
    """Process-wide defaults for reference code that cannot pass `options` (it calls odeint(func, x, t, method=...))."""
    layout = "tbd"
    precision = "fp32"
    ckpt_capacity = 64      # accepted dopri5 steps kept for backprop before GODE_ST_CKPT_OVERFLOW
    log_capacity = 1024
    # Data parallelism: trajectories shard across ranks, weights are replicated, and the flat ODE
    # parameter-gradient buffer the backward kernel writes is all-reduced (sum) in place on the same stream.
    # None: off.  True: the default process group.  A ProcessGroup: that group.
    grad_allreduce = None
    # odeint_adjoint + dopri5: 'continuous' = torchdiffeq's adjoint re-solve; 'discrete' = gradient of the recorded steps
    dopri5_adjoint = "continuous"
    # options={'norm': 'world'}: set by gan_ode_b200.dist.enable_world_norm() (exchange buffers in NVLink peer memory)
    world_norm = None
    # dopri5 backprop: parameter-gradient all-reduce fused into the backward kernel (dist.enable_fused_grad_exchange())
    grad_exchange = None
    # NVTX ranges ("gode.<solver>.fwd" / ".bwd") around the boundary calls, for `ncu --nvtx` / timeline tools; off by default
    nvtx = False
    # Programmatic dependent launch of the rk4 / dopri5 backward kernels (GODE_LAUNCH_PDL_BWD): the backward's weight-staging
    # prologue overlaps the tail of the kernel enqueued just before it on the stream.  Opt-in, because the caller must
    # guarantee that that kernel does not write the ODEFunc weights — true when it is the matching forward (GraphedSolveStep
    # sets it for its own capture), not true in general (e.g. an optimiser step right before a backward).
    pdl = False
    # The thin C++ host (gan_ode_b200._gode_torch) for rk4 / dopri5 calls; False forces the Python autograd.Functions
    use_cpp_host = True


config = _Config()


def _nvtx(name):
    """Wrap an autograd forward/backward in an NVTX range when config.nvtx is set (one attribute test otherwise)."""
    def deco(fn):
        def wrapped(*a, **k):
            if not config.nvtx:
                return fn(*a, **k)
            torch.cuda.nvtx.range_push(name)
            try:
                return fn(*a, **k)
            finally:
                torch.cuda.nvtx.range_pop()
        wrapped.__name__, wrapped.__doc__ = fn.__name__, fn.__doc__
        return wrapped
    return deco


# ------------------------------------------------------------------------------------------------------------
_FORWARD_OK = set()   # types whose call has been probed: obj(t, y) == obj.<mlp>(y) for two values of t


def _has_hooks(*mods):
    for m in mods:
        if getattr(m, "_forward_hooks", None) or getattr(m, "_forward_pre_hooks", None):
            return True
    return False


def verify_call_is_mlp(obj, calls, what):
    """The kernels integrate the MLPs found by STRUCTURE; this checks that the object's own call really is that MLP and
    nothing more (a forward that is time-dependent, negated, scaled or residual around self.fn would otherwise be solved as
    self.fn(x) without an error).  `calls`: [(callable(t, y), mlp module)].  Forward hooks are rejected on every call; the
    numeric probe — two values of t on a fixed 3-row batch, bit-exact comparison, no RNG consumed — runs once per type."""
    mods = [m for _, m in calls]
    if _has_hooks(obj, *mods) or any(_has_hooks(*m._modules.values()) for m in mods):
        raise NotImplementedError("{} has forward hooks registered; the fused kernels would not run them".format(what))
    tp = type(obj)
    if tp in _FORWARD_OK:
        return
    w = calls[0][1][0].weight
    if w.is_cuda and torch.cuda.is_current_stream_capturing():
        return   # cannot synchronise inside a capture: probed at the next eager call
    with torch.no_grad():
        y = torch.linspace(-1.0, 1.0, 3 * w.shape[1], dtype=torch.float32).view(3, w.shape[1]).to(device=w.device, dtype=w.dtype)
        for call, mlp in calls:
            ref = mlp(y)
            for tv in (0.0, 0.73):
                out = call(torch.tensor(tv, dtype=w.dtype).to(w.device), y)
                if not (torch.is_tensor(out) and out.shape == ref.shape and torch.equal(out, ref)):
                    raise NotImplementedError(
                        "{}: calling it does not return its Linear-Tanh-Linear stack applied to x (time-dependent, scaled or "
                        "residual forward?); gan_ode_b200 fuses the solver with the reference's field only, there is no "
                        "generic fallback".format(what))
    _FORWARD_OK.add(tp)


def recognise_field(func):
    """Return (W1, b1, W2, b2) if `func` is the reference's ODEFunc — structure AND forward — else raise NotImplementedError."""
    fn = getattr(func, "fn", None)
    mods = getattr(fn, "_modules", None) if type(fn) is nn.Sequential else None   # fast path: no Sequential.__getitem__
    if mods is not None and len(mods) == 3:
        l0, act, l2 = mods.get("0"), mods.get("1"), mods.get("2")
        if (type(l0) is nn.Linear and type(act) is nn.Tanh and type(l2) is nn.Linear and l0.bias is not None
                and l2.bias is not None and l0.out_features == l2.in_features and l0.in_features == l2.out_features):
            verify_call_is_mlp(func, [(func, fn)], "func ({})".format(type(func).__name__))
            return l0.weight, l0.bias, l2.weight, l2.bias
    if mods is not None and len(mods) == 2:
        # models/mocogan_mnist.py:6-16: f(x) = tanh(W x + b).  Runs on the same kernels as W2 tanh(W1 x + b1) + b2 with
        # W2 = I, b2 = 0 — bit-exact in fp32 (sum of h_d and exact zeros); the two constants take no gradient.
        l0, act = mods.get("0"), mods.get("1")
        if (type(l0) is nn.Linear and type(act) is nn.Tanh and l0.bias is not None and l0.in_features == l0.out_features):
            verify_call_is_mlp(func, [(func, fn)], "func ({})".format(type(func).__name__))
            return (l0.weight, l0.bias) + _identity_layer(l0.in_features, l0.weight.device)
    ok = (isinstance(fn, nn.Sequential) and len(fn) == 3 and isinstance(fn[0], nn.Linear)
          and isinstance(fn[1], nn.Tanh) and isinstance(fn[2], nn.Linear)
          and fn[0].bias is not None and fn[2].bias is not None
          and fn[0].out_features == fn[2].in_features and fn[0].in_features == fn[2].out_features)
    if not ok:
        raise NotImplementedError(
            "gan_ode_b200 fuses the solver with the reference's ODEFunc (models/mocogan_ode.py:6-17): func.fn must be "
            "nn.Sequential(nn.Linear(D,H), nn.Tanh(), nn.Linear(H,D)) — or nn.Sequential(nn.Linear(D,D), nn.Tanh()), "
            "models/mocogan_mnist.py:6-16; got {!r}. There is no generic fallback."
            .format(type(func).__name__))
    verify_call_is_mlp(func, [(func, fn)], "func ({})".format(type(func).__name__))
    return fn[0].weight, fn[0].bias, fn[2].weight, fn[2].bias


_IDENTITY = {}


def _identity_layer(D, device):
    k = (D, str(device))
    v = _IDENTITY.get(k)
    if v is None:
        v = _IDENTITY[k] = (torch.eye(D, dtype=torch.float32, device=device), torch.zeros(D, dtype=torch.float32, device=device))
    return v


def field_parameters(func):
    """The nn.Parameters of a recognised field, in parameters() order (the single-layer field has two)."""
    return [q for q in recognise_field(func) if isinstance(q, nn.Parameter)]


_SIZE_CACHE = {}
_SUPPORTED = set()   # (D, H, precision) triples the library has confirmed


def _rk4_sizes(L, B, D, H, T):
    """(param count, backward workspace bytes) — two ctypes calls, cached per shape."""
    k = (B, D, H, T)
    v = _SIZE_CACHE.get(k)
    if v is None:
        v = (L.gode_param_count(D, H), L.gode_rk4_bwd_workspace_bytes(B, D, H, T))
        if len(_SIZE_CACHE) < 256:
            _SIZE_CACHE[k] = v
    return v


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream() -> int:
    """cudaStream_t of torch's current stream on the current device (the raw getter skips building a Stream object)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


# ---- persistent workspaces (include/gode.h "WORKSPACES ARE PERSISTENT") -----------------------------------------------------
# The front of every workspace is the kernels' grid-synchronisation region; it is zero-filled ONCE and then reused by every
# launch ordered on the same stream, so no memset is enqueued per call.  One workspace per (device, stream) for eager calls; one
# per (device, capture) while a CUDA graph is being captured, taken from a small pool of workspaces that were zero-filled
# ahead of time (a capture cannot run an uncaptured memset; if the pool is empty the workspace is zero-filled INSIDE the capture,
# which is correct and costs one memset node per replay).  Launches that may run concurrently never share a workspace.
_WS_DEFAULT = 8 << 20
_WS_SPARES = 4
_ws_cache = {}    # key -> uint8 tensor
_ws_spare = {}    # device index -> [zero-filled tensors of _WS_DEFAULT bytes]
_cap_id = C.c_ulonglong(0)


def _new_ws(device, nbytes):
    return torch.zeros(max(int(nbytes), _WS_DEFAULT), dtype=torch.uint8, device=device)


def _workspace(device, nbytes):
    """(tensor, pointer) of the persistent workspace of the current stream / the current capture, at least `nbytes` long."""
    L = _lib.lib()
    st = _stream()
    capturing = L.gode_stream_capture_id(st, C.byref(_cap_id)) == 1
    key = (device.index, "cap", _cap_id.value) if capturing else (device.index, st)
    ws = _ws_cache.get(key)
    if ws is not None and ws.numel() >= nbytes:
        return ws
    spares = _ws_spare.setdefault(device.index, [])
    if capturing:
        ws = spares.pop() if (spares and nbytes <= _WS_DEFAULT) else _new_ws(device, nbytes)   # else: memset node in the graph
    else:
        ws = _new_ws(device, nbytes)
        while len(spares) < _WS_SPARES:
            spares.append(_new_ws(device, _WS_DEFAULT))
        if len(_ws_cache) > 256:      # streams come and go: forget eager entries (captured graphs keep theirs alive below)
            for k in [k for k in _ws_cache if len(k) == 2]:
                del _ws_cache[k]
    _ws_cache[key] = ws
    return ws


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _f32c(t: torch.Tensor) -> torch.Tensor:
    """fp32, contiguous, 16-byte aligned (kernels use 128-bit accesses).  The common case returns `t` itself."""
    if t.dtype is torch.float32 and t.is_contiguous() and not (t.data_ptr() & 15):
        return t
    t = t.detach().float().contiguous()
    if t.data_ptr() & 15:
        t = t.clone()
    return t


def _require_cuda(*tensors, what="y0", weights=()):
    """The one device gate of the host side: the product path runs on the GPU or not at all, and every pointer that
    crosses the C ABI must belong to ONE device (a CPU-resident module, or weights on another GPU, would otherwise reach a
    kernel as a raw host / foreign pointer and fault the whole CUDA context instead of raising here)."""
    for x in tensors:
        if not x.is_cuda:
            raise GodeError("{} is on {}: the B200 path has no CPU fallback".format(what, x.device))
    if tensors:
        dev = tensors[0].device
        for x in tensors[1:]:
            if x.device != dev:
                raise GodeError("{}: tensors are on different devices ({} and {})".format(what, dev, x.device))
        for w in weights:
            if w.device != dev:
                raise GodeError("the field's parameters are on {} but {} is on {}: move the module to the solve's device "
                                "(there is no CPU fallback and no cross-device path)".format(w.device, what, dev))


class _on_device:
    """Make `device` the current CUDA device around a C-ABI call (the library launches on the current device and
    `_stream()` returns the current device's current stream).  The common case — already current — does nothing."""
    __slots__ = ("idx", "prev")

    def __init__(self, device):
        self.idx, self.prev = device.index, None

    def __enter__(self):
        if self.idx is None:       # not a CUDA tensor: the device gate has already decided what happens
            return
        cur = torch.cuda.current_device()
        if cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
        return False


def _bwd_on_device(fn):
    """autograd.Function.backward under the device of the incoming gradient (autograd worker threads normally have it set
    already; a backward driven from another thread or another current device must not launch on the wrong GPU)."""
    def wrapped(ctx, grad, *rest):
        with _on_device(grad.device):
            return fn(ctx, grad, *rest)
    wrapped.__name__, wrapped.__doc__ = fn.__name__, fn.__doc__
    return wrapped


def _check_common(y0, t):
    if not isinstance(y0, torch.Tensor):
        raise NotImplementedError("tuple-valued y0 is not on the gan-ode hot path")
    if not torch.is_floating_point(y0):
        raise TypeError("`y0` must be a floating point Tensor but is a {}".format(y0.type()))
    if not isinstance(t, torch.Tensor):
        raise TypeError("`t` must be a torch.Tensor")
    assert t.ndimension() == 1, "t must be one dimensional"
    if not torch.is_floating_point(t):
        raise TypeError("`t` must be a floating point Tensor but is a {}".format(t.type()))
    if y0.dtype != torch.float32:
        raise NotImplementedError("the fused path computes in float32 state (reference dtype); got {}".format(y0.dtype))
    if y0.dim() != 2:
        raise NotImplementedError("y0 must be (B, D) as in models/mocogan_ode.py:136-140; got {}".format(tuple(y0.shape)))
    _require_cuda(y0)
    if len(t) < 2:
        raise NotImplementedError("len(t) must be >= 2")


def _host_steps(t: torch.Tensor):
    """t on the host (the reference's case: torch.linspace(...) on CPU, models/mocogan_ode.py:143) -> numpy views.
    Returns (t64_increasing, dt32_signed, fsign).  torchdiffeq _check_inputs: t must be strictly monotone; a
    decreasing grid is integrated as increasing -t with the field negated (fsign = -1).  dt_j = t[j+1]-t[j] is
    taken in t's own dtype and then rounded to fp32, which is what multiplying it into fp32 state does upstream."""
    tn = t.detach().numpy()
    key = (tn.dtype.str, tn.tobytes())
    hit = _HOST_STEPS_CACHE.get(key)
    if hit is not None:
        return hit
    out = _host_steps_uncached(tn)
    if len(_HOST_STEPS_CACHE) < 64:
        _HOST_STEPS_CACHE[key] = out
    return out


_HOST_STEPS_CACHE = {}   # immutable values (numpy arrays are never written after creation); bounded


def _host_steps_uncached(tn):
    d = np.diff(tn)
    if (d > 0).all():
        fsign = 1.0
    elif (d < 0).all():
        fsign = -1.0
    else:
        raise AssertionError("t must be strictly increasing or decreasing")
    t64 = tn.astype(np.float64)
    if fsign < 0:
        t64 = -t64
    return t64, np.ascontiguousarray(d, dtype=np.float32), fsign


def _layout_code(name):
    if name not in ("tbd", "btd"):
        raise ValueError("options['layout'] must be 'tbd' or 'btd'")
    return _lib.LAYOUT_TBD if name == "tbd" else _lib.LAYOUT_BTD


def _alloc_traj(T, B, D, layout, like):
    if layout == _lib.LAYOUT_TBD:
        buf = torch.empty((T, B, D), dtype=torch.float32, device=like.device)
        return buf, buf
    buf = torch.empty((B, T, D), dtype=torch.float32, device=like.device)
    return buf, buf.transpose(0, 1)


def _grad_in_layout(g: torch.Tensor, layout):
    """Upstream gradient as a contiguous buffer in the kernel's layout (zero-copy when it already is)."""
    if layout == _lib.LAYOUT_TBD and g.dtype is torch.float32 and g.is_contiguous() and not (g.data_ptr() & 15):
        return g   # the common case (autograd hands a fresh contiguous tensor): no ops at all
    g = g.detach()
    if g.dtype != torch.float32:
        g = g.float()
    if layout == _lib.LAYOUT_BTD:
        g = g.transpose(0, 1)
    g = g.contiguous()
    if g.data_ptr() % 16:
        g = g.clone()
    return g


def _maybe_allreduce(flat):
    g = config.grad_allreduce
    if g is None or g is False:
        return
    if callable(g):          # gan_ode_b200.dist.P2PAllReduce: fused one-shot kernel over NVLink peer memory
        if flat.numel() <= g.small and not (flat.data_ptr() & 15):  # larger vectors: NCCL's multi-channel ring wins
            g(flat)
            return
        g = True
    import torch.distributed as dist
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=None if g is True else g)


def _split_params(flat, D, H, needs, reduced=False):
    if not reduced:
        _maybe_allreduce(flat)
    n1 = H * D
    outs = (flat[:n1].view(H, D), flat[n1:n1 + H], flat[n1 + H:n1 + H + D * H].view(D, H), flat[n1 + H + D * H:])
    return tuple(o if need else None for o, need in zip(outs, needs))


# ------------------------------------------------------------------------------------------------------------
# fixed-grid rk4
class _Rk4(torch.autograd.Function):
    """forward: gode_rk4_fwd (one launch).  backward: gode_rk4_adjoint_bwd (torchdiffeq odeint_adjoint semantics) or
    gode_rk4_backprop_bwd (autograd-through-odeint semantics), one launch."""

    @staticmethod
    @_nvtx("gode.rk4.fwd")
    def forward(ctx, y0, dt, meta, W1, b1, W2, b2):
        L = _lib.lib()
        B, D = y0.shape
        H = W1.shape[0]
        T = meta["T"]
        y0c, W1c, b1c, W2c, b2c = (_f32c(x) for x in (y0, W1, b1, W2, b2))
        buf, view = _alloc_traj(T, B, D, meta["layout"], y0)
        dt_ptr, dt_dev = _dt_arg(dt)
        if meta.get("method", 0):   # euler / midpoint (FP32, D = H = 16): same kernels, other tableau
            rc = L.gode_fixed_fwd(meta["method"], y0c.data_ptr(), W1c.data_ptr(), b1c.data_ptr(), W2c.data_ptr(),
                                  b2c.data_ptr(), dt_ptr, dt_dev, B, D, H, T, meta["layout"], buf.data_ptr(), _stream())
        else:
            rc = L.gode_rk4_fwd(y0c.data_ptr(), W1c.data_ptr(), b1c.data_ptr(), W2c.data_ptr(), b2c.data_ptr(), dt_ptr,
                                dt_dev, B, D, H, T, meta["precision"], meta["layout"], buf.data_ptr(), _stream())
        if rc:
            _lib.check(rc, "gode_rk4_fwd")
        ctx.meta = meta
        ctx.dt = dt
        ctx.save_for_backward(buf, W1c, b1c, W2c, b2c)
        return view

    @staticmethod
    @_nvtx("gode.rk4.bwd")
    @_bwd_on_device
    def backward(ctx, grad_traj):
        L = _lib.lib()
        buf, W1c, b1c, W2c, b2c = ctx.saved_tensors
        meta, dt = ctx.meta, ctx.dt
        T = meta["T"]
        if meta["layout"] == _lib.LAYOUT_TBD:
            _, B, D = buf.shape
        else:
            B, _, D = buf.shape
        H = W1c.shape[0]
        g = _grad_in_layout(grad_traj, meta["layout"])
        n_param, ws_bytes = _rk4_sizes(L, B, D, H, T)
        # one allocation for both gradient outputs (grad_p first: it stays 16-byte aligned for the all-reduce kernels)
        n_pad = (n_param + 3) & ~3
        gbuf = torch.empty(n_pad + B * D, dtype=torch.float32, device=buf.device)
        grad_p, grad_y0 = gbuf[:n_param], gbuf[n_pad:].view(B, D)
        ws = _workspace(buf.device, ws_bytes)
        fn = L.gode_rk4_adjoint_bwd if meta["adjoint"] else L.gode_rk4_backprop_bwd
        dt_ptr, dt_dev = _dt_arg(dt)
        bwd_prec = _bwd_precision(meta, D, H)
        ex = config.grad_exchange
        fused = (ex is not None and not meta.get("method", 0) and bwd_prec == _lib.PREC["fp32"] and (D, H) == (16, 16))
        pdl = config.pdl
        if pdl:
            L.gode_set_thread_launch_flags(_lib.LAUNCH_PDL_BWD)
        if fused:    # all-reduce over ranks inside the kernel's reduction tail (NVLink peer memory)
            xs = ex.struct(n_param)
            rc = L.gode_rk4_bwd_world(1 if meta["adjoint"] else 0, buf.data_ptr(), g.data_ptr(), W1c.data_ptr(), b1c.data_ptr(),
                                      W2c.data_ptr(), b2c.data_ptr(), dt_ptr, dt_dev, B, D, H, T, meta["layout"],
                                      grad_y0.data_ptr(), grad_p.data_ptr(), ws.data_ptr(), ws_bytes, C.byref(xs), _stream())
        elif meta.get("method", 0):
            fn = L.gode_fixed_adjoint_bwd if meta["adjoint"] else L.gode_fixed_backprop_bwd
            rc = fn(meta["method"], buf.data_ptr(), g.data_ptr(), W1c.data_ptr(), b1c.data_ptr(), W2c.data_ptr(),
                    b2c.data_ptr(), dt_ptr, dt_dev, B, D, H, T, meta["layout"], grad_y0.data_ptr(), grad_p.data_ptr(),
                    ws.data_ptr(), ws_bytes, _stream())
        else:
            rc = fn(buf.data_ptr(), g.data_ptr(), W1c.data_ptr(), b1c.data_ptr(), W2c.data_ptr(), b2c.data_ptr(), dt_ptr,
                    dt_dev, B, D, H, T, bwd_prec, meta["layout"], grad_y0.data_ptr(), grad_p.data_ptr(),
                    ws.data_ptr(), ws_bytes, _stream())
        if pdl:
            L.gode_set_thread_launch_flags(0)
        if rc:
            _lib.check(rc, "gode_rk4_bwd")
        needs = ctx.needs_input_grad
        gW1, gb1, gW2, gb2 = _split_params(grad_p, D, H, needs[3:7], reduced=fused)
        return (grad_y0 if needs[0] else None), None, None, gW1, gb1, gW2, gb2


def _bwd_precision(meta, D, H):
    """Which adjoint kernel backs a tensor-core forward: wide field (D=64, H=256) in bf16 mode -> the tcgen05 adjoint
    (csrc/tc_rk4_adj_wide.cu) by default; reference shape -> opt-in (options['bwd_precision']='bf16', tc_rk4_adj_small.cu).
    Everything else backpropagates with the FP32 kernels from the stored trajectory (inside the 2e-3 budget of tf32/bf16)."""
    want = meta.get("bwd_precision")
    if meta["adjoint"] and meta["precision"] == _lib.PREC["bf16"]:
        if (D, H) == (64, 256) and want in (None, "bf16"):
            return _lib.PREC["bf16"]
        if (D, H) == (16, 16) and want == "bf16":
            return _lib.PREC["bf16"]
    return _lib.PREC["fp32"]


def _dt_arg(dt):
    """(pointer, on_device) of a step table held as a host numpy array or a device tensor."""
    if isinstance(dt, np.ndarray):
        return dt.ctypes.data, 0
    return dt.data_ptr(), 1


def _substepped(apply_fixed, y0, t, step_size, layout_name):
    """options={'step_size': h} of torchdiffeq's fixed-grid solvers (solvers.py::FixedGridODESolver): integrate on the grid
    t0, t0+h, t0+2h, ... (last point clamped to t[-1]) and LINEARLY interpolate the requested times between grid states
    (`_linear_interp`; a requested time that IS a grid point returns that state exactly).  The fine-grid solve is one
    kernel launch; the interpolation is a handful of differentiable torch ops, so gradients reach the kernel's backward as
    an upstream gradient on the fine grid.  Offered for `odeint` (backprop-through-solver), where this is exact; the
    adjoint variant re-solves sub-stepped intervals without resetting the state at the inner grid points and is not built."""
    tc = t.detach().cpu()
    sign = 1.0 if bool(tc[-1] > tc[0]) else -1.0
    ts = tc * sign                                   # torchdiffeq integrates a decreasing grid as increasing -t
    h = torch.as_tensor(step_size, dtype=ts.dtype)
    niters = int(torch.ceil((ts[-1] - ts[0]) / h + 1).item())
    grid = torch.arange(0, niters, dtype=ts.dtype) * h + ts[0]
    grid[-1] = ts[-1]
    fine = apply_fixed(grid * sign)                  # (len(grid), B, D), always time-major here
    outs = [fine[0]]
    k = 0
    for j in range(1, len(ts)):
        while not bool(grid[k + 1] >= ts[j]):
            k += 1
        t0, t1, tj = grid[k], grid[k + 1], ts[j]
        if bool(tj == t1):
            outs.append(fine[k + 1])
        elif bool(tj == t0):
            outs.append(fine[k])
        else:
            slope = float((tj - t0) / (t1 - t0))
            outs.append(fine[k] + slope * (fine[k + 1] - fine[k]))
    sol = torch.stack(outs, 0)
    if layout_name == "btd":
        sol = sol.transpose(0, 1).contiguous().transpose(0, 1)
    return sol


_SUBSTEP_CACHE = {}


def _adjoint_substeps(t: torch.Tensor, step_size, device):
    """Sub-step tables of torchdiffeq's adjoint when options['step_size'] is set: adjoint.py solves every output interval with
    odeint(aug, state, [t_i, t_{i-1}], method, options) and FixedGridODESolver lays ITS grid from the interval's own start:
    t_i, t_i -+ h, ... with the last point clamped to t_{i-1} (a decreasing pair is integrated as increasing negated time).
    Returns device tensors (sub_dt float32, signed like the kernel's dt; sub_beg, sub_end int32 indexed by i)."""
    tc = t.detach().cpu()
    key = (tc.dtype, tc.numpy().tobytes(), float(step_size), str(device))
    hit = _SUBSTEP_CACHE.get(key)
    if hit is not None:
        return hit
    T = len(tc)
    h = torch.as_tensor(step_size, dtype=tc.dtype)
    dts, beg, end = [], [0] * T, [0] * T
    for i in range(T - 1, 0, -1):
        a, b = tc[i], tc[i - 1]
        rev = bool(a > b)                       # forward grid increasing: the reversed solve runs in negated time
        s0, s1 = (-a, -b) if rev else (a, b)
        niters = int(torch.ceil((s1 - s0) / h + 1).item())
        g = torch.arange(0, niters, dtype=tc.dtype) * h + s0
        g[-1] = s1
        d = (g[1:] - g[:-1]).to(torch.float32) * (1.0 if rev else -1.0)
        beg[i] = len(dts)
        dts.extend(d.tolist())
        end[i] = len(dts)
    out = (torch.tensor(dts, dtype=torch.float32).to(device), torch.tensor(beg, dtype=torch.int32).to(device),
           torch.tensor(end, dtype=torch.int32).to(device))
    if len(_SUBSTEP_CACHE) < 64:
        _SUBSTEP_CACHE[key] = out
    return out


class _FixedSubstepAdjoint(torch.autograd.Function):
    """odeint_adjoint with options['step_size'] on a fixed-grid method: forward = fine-grid kernel solve + linear
    interpolation of the requested times (no graph), backward = gode_fixed_adjoint_bwd_substep."""

    @staticmethod
    def forward(ctx, y0, meta, W1, b1, W2, b2):
        inner = dict(meta, adjoint=False)

        def apply_fixed(grid):
            m = dict(inner, T=len(grid), layout=_lib.LAYOUT_TBD)
            return _Rk4.apply(y0.detach(), _rk4_dt(grid, {}, y0.device), m, W1.detach(), b1.detach(), W2.detach(), b2.detach())

        with torch.no_grad():
            sol = _substepped(apply_fixed, y0, meta["t"], meta["step_size"], "tbd")      # (T, B, D)
        if meta["layout"] == _lib.LAYOUT_TBD:
            buf = sol.contiguous()
            view = buf
        else:
            buf = sol.transpose(0, 1).contiguous()
            view = buf.transpose(0, 1)
        ctx.meta = meta
        ctx.save_for_backward(buf, *(_f32c(x) for x in (W1, b1, W2, b2)))
        return view

    @staticmethod
    @_bwd_on_device
    def backward(ctx, grad_traj):
        L = _lib.lib()
        buf, W1c, b1c, W2c, b2c = ctx.saved_tensors
        meta = ctx.meta
        T = meta["T"]
        B, D = (buf.shape[1], buf.shape[2]) if meta["layout"] == _lib.LAYOUT_TBD else (buf.shape[0], buf.shape[2])
        H = W1c.shape[0]
        g = _grad_in_layout(grad_traj, meta["layout"])
        n_param, ws_bytes = _rk4_sizes(L, B, D, H, T)
        grad_p = torch.empty(n_param, dtype=torch.float32, device=buf.device)
        grad_y0 = torch.empty((B, D), dtype=torch.float32, device=buf.device)
        ws = _workspace(buf.device, ws_bytes)
        sub_dt, sub_beg, sub_end = _adjoint_substeps(meta["t"], meta["step_size"], buf.device)
        rc = L.gode_fixed_adjoint_bwd_substep(meta.get("method", 0), buf.data_ptr(), g.data_ptr(), W1c.data_ptr(), b1c.data_ptr(),
                                              W2c.data_ptr(), b2c.data_ptr(), sub_dt.data_ptr(), sub_beg.data_ptr(),
                                              sub_end.data_ptr(), B, D, H, T, meta["layout"], grad_y0.data_ptr(),
                                              grad_p.data_ptr(), ws.data_ptr(), ws_bytes, _stream())
        _lib.check(rc, "gode_fixed_adjoint_bwd_substep")
        needs = ctx.needs_input_grad
        gW1, gb1, gW2, gb2 = _split_params(grad_p, D, H, needs[2:6])
        return (grad_y0 if needs[0] else None), None, gW1, gb1, gW2, gb2


def _rk4_dt(t: torch.Tensor, options, device):
    """Step table of torchdiffeq's fixed-grid driver: grid == t when no step_size is given, dt_j = t[j+1]-t[j] in t's
    dtype, multiplied into fp32 state (=> rounded to fp32).  Decreasing t gives negative dt, which is bit-identical
    to upstream's (-t, -f) rewrite for the 3/8 rule (negation is exact).  A host t gives a host table that rides in
    the kernel launch parameters (no copy, no sync); a device t stays on the device (monotonicity then unchecked
    unless options['check'])."""
    if options.get("grid_constructor") is not None:
        raise NotImplementedError("grid_constructor is not supported (options['step_size'] is, for odeint)")
    if not t.is_cuda:
        _, dt, _ = _host_steps(t)
        if len(dt) > _lib.MAX_HOST_STEPS:
            return torch.from_numpy(dt).to(device)
        return dt
    d = t[1:] - t[:-1]
    if options.get("check", False):
        assert bool((d > 0).all()) or bool((d < 0).all()), "t must be strictly increasing or decreasing"
    return d.to(torch.float32).contiguous()


# ------------------------------------------------------------------------------------------------------------
# adaptive dopri5
class StepLog:
    """Host view of a device-side GodeStepLog (+ attempt arrays).  Reading any attribute synchronises."""

    def __init__(self, raw: torch.Tensor, cap: int):
        self._raw, self._cap, self._host = raw, cap, None

    def _load(self):
        if self._host is None:
            h = self._raw.cpu().numpy().tobytes()
            hdr = GodeStepLog.from_buffer_copy(h[:C.sizeof(GodeStepLog)])
            import numpy as np
            cap = self._cap
            off = 64
            n = min(hdr.n_attempts, cap)
            t0 = np.frombuffer(h, dtype=np.float64, count=cap, offset=off)[:n]
            dt = np.frombuffer(h, dtype=np.float64, count=cap, offset=off + 8 * cap)[:n]
            er = np.frombuffer(h, dtype=np.float32, count=cap, offset=off + 16 * cap)[:n]
            acc = np.frombuffer(h, dtype=np.uint8, count=cap, offset=off + 20 * cap)[:n]
            self._host = dict(status=hdr.status, n_attempts=hdr.n_attempts, n_accepted=hdr.n_accepted, nfe=hdr.nfe,
                              dt0=hdr.dt0, t_final=hdr.t_final, t0=t0.tolist(), dt=dt.tolist(),
                              error_ratio=er.tolist(), accepted=[bool(a) for a in acc])
        return self._host

    def __getattr__(self, k):
        if k.startswith("_"):
            raise AttributeError(k)
        return self._load()[k]

    @property
    def n_rejected(self):
        return self.n_attempts - self.n_accepted


_LAST_LOG = [None]


def last_step_log() -> Optional[StepLog]:
    """Step log of the most recent adaptive solve on this process (synchronises when read)."""
    return _LAST_LOG[0]


# ---- status mailbox: failed adaptive solves are loud without a synchronisation per call ----------------------------------------
# One int in pinned (device-mapped) host memory, registered with the library once.  A solve that ends with a non-zero status
# word also stores it there (nothing is written on the success path).  The host looks at it — a plain memory read — at every
# entry into the solver API, in the backward of the adaptive solves, and at the natural synchronisation points
# (StepLog reads, GraphedSolveStep.sync, gan_ode_b200.check_status()).  So a failed forward is reported at the latest by the
# next call into this package, instead of only poisoning the gradients with NaN.
_mailbox = [None, None]   # (pinned tensor, numpy view)


def _mailbox_view():
    mb = _mailbox[1]
    if mb is None and _mailbox[0] is None and torch.cuda.is_available():
        t = torch.zeros(16, dtype=torch.int32).pin_memory()
        _lib.lib().gode_set_status_mailbox(t.data_ptr())   # UVA: pinned host memory is addressable from every device as is
        _mailbox[0], _mailbox[1] = t, t.numpy()
        mb = _mailbox[1]
    return mb


def check_status():
    """Raise (torchdiffeq's assert texts) if any adaptive solve launched so far has ended with a solver failure that has not
    been reported yet.  Never synchronises: call it after a synchronisation of your own to be sure nothing is in flight."""
    mb = _mailbox[1]
    if mb is not None and mb[0]:
        st = int(mb[0])
        mb[0] = 0
        try:
            raise_for_status(st)
        except (AssertionError, GodeError) as e:
            raise type(e)("{} (reported by the device for an earlier adaptive solve; gradients computed from it are NaN)"
                          .format(e)) from None


def raise_for_status(status: int):
    """torchdiffeq's solver asserts, raised from the device status word.  (Callers read `status` from a step log, i.e. after a
    synchronisation: whatever the mailbox holds by then has been reported here, so it is cleared.)"""
    if status and _mailbox[1] is not None:
        _mailbox[1][0] = 0
    if status & _lib.ST_NONFINITE:
        raise AssertionError("non-finite values in state `y`")
    if status & _lib.ST_DT_UNDERFLOW:
        raise AssertionError("underflow in dt")
    if status & _lib.ST_MAX_STEPS:
        raise AssertionError("max_num_steps exceeded")
    if status & _lib.ST_PEER_TIMEOUT:
        raise GodeError("world-scope norm: a peer rank did not publish its partial sum within 10 s")
    if status & _lib.ST_CKPT_OVERFLOW:
        raise GodeError("more accepted steps than options['ckpt_capacity']; raise gan_ode_b200.config.ckpt_capacity")


def _log_layout(cap):
    # [GodeStepLog (64 B slot)] [att_t0: cap f64] [att_dt: cap f64] [att_er: cap f32] [att_acc: cap u8]
    return 64 + 8 * cap + 8 * cap + 4 * cap + cap


class _Dopri5(torch.autograd.Function):
    """forward: gode_dopri5_fwd (one cooperative launch, batch-global error norm).
    backward: gode_dopri5_backprop_bwd — reverse-mode through the accepted steps recorded on the device."""

    @staticmethod
    @_nvtx("gode.dopri5.fwd")
    def forward(ctx, y0, meta, W1, b1, W2, b2):
        L = _lib.lib()
        B, D = y0.shape
        H = W1.shape[0]
        T = meta["T"]
        dev = y0.device
        y0c, W1c, b1c, W2c, b2c = (_f32c(x) for x in (y0, W1, b1, W2, b2))
        buf, view = _alloc_traj(T, B, D, meta["layout"], y0)
        o = meta["opts"]
        cap = o.log_capacity
        raw = torch.empty(_log_layout(cap), dtype=torch.uint8, device=dev)
        base = raw.data_ptr()
        keep = meta["keep_ckpt"]
        kc = o.ckpt_capacity if keep else 0
        opts = GodeAdaptiveOpts.from_buffer_copy(bytes(o))
        opts.ckpt_capacity = kc
        # adaptive_heun is not FSAL: the forward checkpoints every step's f0 beside its y0 (second half of the buffer)
        rows = max(kc, 1) * (2 if o.tableau == _lib.TABLEAUS["adaptive_heun"] else 1)
        ckpt = torch.empty((rows, B, D), dtype=torch.float32, device=dev) if keep else None
        acc = torch.empty(2 * max(kc, 1), dtype=torch.float64, device=dev) if keep else None
        ws_bytes = L.gode_dopri5_workspace_bytes(B, D, H)
        ws = _workspace(dev, ws_bytes)
        tarr = meta["t64"]
        args = (_ptr(y0c), _ptr(W1c), _ptr(b1c), _ptr(W2c), _ptr(b2c), tarr.ctypes.data, B, D, H, T, C.byref(opts),
                meta["layout"], _ptr(buf), base, base + 64, base + 64 + 8 * cap, base + 64 + 16 * cap, base + 64 + 20 * cap,
                _ptr(ckpt), _ptr(acc), (acc.data_ptr() + 8 * max(kc, 1)) if keep else None, _ptr(ws), ws_bytes)
        wn = meta.get("world")
        if wn is not None:   # error norm over the trajectories of all ranks, exchanged inside the kernel over peer memory
            wstruct = wn.struct(B)
            _lib.check(L.gode_dopri5_fwd_world(*args, C.byref(wstruct), _stream()), "gode_dopri5_fwd_world")
        else:
            _lib.check(L.gode_dopri5_fwd(*args, _stream()), "gode_dopri5_fwd")
        log = StepLog(raw, cap)
        _LAST_LOG[0] = log
        if meta["check"]:
            raise_for_status(log.status)
        ctx.meta, ctx.log, ctx.kc, ctx.tarr = meta, log, kc, tarr
        ctx.save_for_backward(raw, ckpt, acc, W1c, b1c, W2c, b2c)
        return view

    @staticmethod
    @_nvtx("gode.dopri5.bwd")
    @_bwd_on_device
    def backward(ctx, grad_traj):
        L = _lib.lib()
        raw, ckpt, acc, W1c, b1c, W2c, b2c = ctx.saved_tensors
        meta, kc = ctx.meta, ctx.kc
        if ckpt is None:
            raise GodeError("dopri5 forward ran without checkpoints (inputs did not require grad)")
        # No host sync by default (keeps fwd+bwd CUDA-graph capturable): a failed forward (status != 0, including
        # checkpoint overflow) makes the backward kernel return NaN gradients; options['check'] raises eagerly, and the
        # status mailbox raises here if the forward has already finished, else at the next call into the package.
        if meta["check"]:
            raise_for_status(ctx.log.status)
        check_status()
        T = meta["T"]
        _, B, D = ckpt.shape
        H = W1c.shape[0]
        g = _grad_in_layout(grad_traj, meta["layout"])
        grad_y0 = torch.empty((B, D), dtype=torch.float32, device=ckpt.device)
        grad_p = torch.empty(L.gode_param_count(D, H), dtype=torch.float32, device=ckpt.device)
        ws_bytes = L.gode_dopri5_backprop_workspace_bytes(B, D, H, kc)   # wide fields: + the gradient rows of kc steps
        ws = _workspace(ckpt.device, ws_bytes)
        args = (_ptr(g), _ptr(W1c), _ptr(b1c), _ptr(W2c), _ptr(b2c), ctx.tarr.ctypes.data, B, D, H, T,
                meta["layout"], raw.data_ptr(), _ptr(ckpt), _ptr(acc), acc.data_ptr() + 8 * kc, kc,
                C.c_float(meta["opts"].fsign), _ptr(grad_y0), _ptr(grad_p), _ptr(ws), ws_bytes)
        ex = config.grad_exchange
        pdl = config.pdl
        if pdl:
            L.gode_set_thread_launch_flags(_lib.LAUNCH_PDL_BWD)
        tab = meta["opts"].tableau
        if tab:              # bosh3 / adaptive_heun: same replay kernel, other tableau
            rc = L.gode_adaptive_backprop_bwd(tab, *args, _stream())
            ex = None
        elif ex is not None:   # the all-reduce over ranks happens inside the kernel's reduction tail (NVLink peer memory)
            xs = ex.struct(grad_p.numel())
            rc = L.gode_dopri5_backprop_bwd_world(*args, C.byref(xs), _stream())
        else:
            rc = L.gode_dopri5_backprop_bwd(*args, _stream())
        if pdl:
            L.gode_set_thread_launch_flags(0)
        if rc:
            _lib.check(rc, "gode_dopri5_backprop_bwd")
        needs = ctx.needs_input_grad
        gW1, gb1, gW2, gb2 = _split_params(grad_p, D, H, needs[2:6], reduced=ex is not None)
        return (grad_y0 if needs[0] else None), None, gW1, gb1, gW2, gb2


_LAST_ADJ_LOG = [None]


def last_adjoint_log() -> Optional[StepLog]:
    """Step log of the most recent continuous dopri5 adjoint solve (attempts of all intervals back to back; `t0` is not
    recorded).  Reading synchronises."""
    v = _LAST_ADJ_LOG[0]
    if isinstance(v, tuple):    # the C++ host ran the backward: fetch the log it kept
        raw = _ext[0].last_adjoint_log()
        return StepLog(raw, v[1]) if raw.numel() else None
    return v


class _Dopri5Adjoint(torch.autograd.Function):
    """odeint_adjoint with the adaptive solver as torchdiffeq computes it: forward gode_dopri5_fwd (nothing kept but the
    outputs), backward gode_dopri5_adjoint_bwd — per output interval a fresh dopri5 solve of (y, a, theta_bar) backwards in
    time under the default adjoint norm (adjoint.py)."""

    @staticmethod
    @_nvtx("gode.dopri5_adjoint.fwd")
    def forward(ctx, y0, meta, W1, b1, W2, b2):
        L = _lib.lib()
        B, D = y0.shape
        H = W1.shape[0]
        T = meta["T"]
        dev = y0.device
        y0c, W1c, b1c, W2c, b2c = (_f32c(x) for x in (y0, W1, b1, W2, b2))
        buf, view = _alloc_traj(T, B, D, meta["layout"], y0)
        o = meta["opts"]
        cap = o.log_capacity
        raw = torch.empty(_log_layout(cap), dtype=torch.uint8, device=dev)
        base = raw.data_ptr()
        opts = GodeAdaptiveOpts.from_buffer_copy(bytes(o))
        opts.ckpt_capacity = 0
        ws_bytes = L.gode_dopri5_workspace_bytes(B, D, H)
        ws = _workspace(dev, ws_bytes)
        tarr = meta["t64"]
        _lib.check(L.gode_dopri5_fwd(
            _ptr(y0c), _ptr(W1c), _ptr(b1c), _ptr(W2c), _ptr(b2c), tarr.ctypes.data, B, D, H, T, C.byref(opts),
            meta["layout"], _ptr(buf), base, base + 64, base + 64 + 8 * cap, base + 64 + 16 * cap, base + 64 + 20 * cap,
            None, None, None, _ptr(ws), ws_bytes, _stream()), "gode_dopri5_fwd")
        log = StepLog(raw, cap)
        _LAST_LOG[0] = log
        if meta["check"]:
            raise_for_status(log.status)
        ctx.meta, ctx.tarr = meta, tarr
        ctx.save_for_backward(buf, W1c, b1c, W2c, b2c)
        return view

    @staticmethod
    @_nvtx("gode.dopri5_adjoint.bwd")
    @_bwd_on_device
    def backward(ctx, grad_traj):
        L = _lib.lib()
        buf, W1c, b1c, W2c, b2c = ctx.saved_tensors
        meta = ctx.meta
        T = meta["T"]
        B, D = (buf.shape[1], buf.shape[2]) if meta["layout"] == _lib.LAYOUT_TBD else (buf.shape[0], buf.shape[2])
        H = W1c.shape[0]
        dev = buf.device
        g = _grad_in_layout(grad_traj, meta["layout"])
        grad_y0 = torch.empty((B, D), dtype=torch.float32, device=dev)
        grad_p = torch.empty(L.gode_param_count(D, H), dtype=torch.float32, device=dev)
        ao = meta["adj_opts"]
        cap = ao.log_capacity
        raw = torch.zeros(_log_layout(cap), dtype=torch.uint8, device=dev)
        base = raw.data_ptr()
        ws_bytes = L.gode_dopri5_adjoint_workspace_bytes(B, D, H)
        ws = _workspace(dev, ws_bytes)
        rc = L.gode_dopri5_adjoint_bwd(
            _ptr(buf), _ptr(g), _ptr(W1c), _ptr(b1c), _ptr(W2c), _ptr(b2c), ctx.tarr.ctypes.data, B, D, H, T,
            meta["layout"], C.byref(ao), meta["param_mask"], _ptr(grad_y0), _ptr(grad_p), base, base + 64 + 8 * cap, base + 64 + 16 * cap,
            base + 64 + 20 * cap, _ptr(ws), ws_bytes, _stream())
        if rc == _lib.ERR_COOP:
            raise GodeError("the continuous dopri5 adjoint keeps the whole batch co-resident (at most 18944 trajectories per "
                            "GPU); shard the batch, or pass options={'adjoint': 'discrete'} for the gradient of the recorded "
                            "steps (gode_dopri5_backprop_bwd)")
        _lib.check(rc, "gode_dopri5_adjoint_bwd")
        log = StepLog(raw, cap)
        _LAST_ADJ_LOG[0] = log
        if meta["check"]:
            raise_for_status(log.status)
        needs = ctx.needs_input_grad
        gW1, gb1, gW2, gb2 = _split_params(grad_p, D, H, needs[2:6])
        return (grad_y0 if needs[0] else None), None, gW1, gb1, gW2, gb2


class TrajStepLog:
    """Host view of a per-trajectory dopri5 solve: `n_accepted`, `n_attempts` are (B,) arrays; `dt`, `error_ratio`,
    `accepted` are (n_max, B) arrays of per-attempt records (rows beyond a trajectory's own n_attempts are undefined).
    Reading any attribute synchronises."""

    def __init__(self, hdr, n_acc, n_att, att, cap, B):
        self._t = (hdr, n_acc, n_att, att, cap, B)
        self._host = None

    def _load(self):
        if self._host is None:
            hdr, n_acc, n_att, att, cap, B = self._t
            h = GodeStepLog.from_buffer_copy(hdr.cpu().numpy().tobytes()[:C.sizeof(GodeStepLog)])
            na, nt = n_acc.cpu().numpy(), n_att.cpu().numpy()
            d = dict(status=h.status, max_attempts=h.n_attempts, max_accepted=h.n_accepted, nfe=h.nfe,
                     n_accepted=na, n_attempts=nt)
            if att is not None:
                raw = att.cpu().numpy().tobytes()
                n = min(int(nt.max()) if len(nt) else 0, cap)
                d["dt"] = np.frombuffer(raw, dtype=np.float64, count=cap * B).reshape(cap, B)[:n]
                d["error_ratio"] = np.frombuffer(raw, dtype=np.float32, count=cap * B, offset=8 * cap * B).reshape(cap, B)[:n]
                d["accepted"] = np.frombuffer(raw, dtype=np.uint8, count=cap * B, offset=12 * cap * B).reshape(cap, B)[:n].astype(bool)
            self._host = d
        return self._host

    def __getattr__(self, k):
        if k.startswith("_"):
            raise AttributeError(k)
        return self._load()[k]


class _Dopri5Traj(torch.autograd.Function):
    """dopri5 with per-trajectory step control (options={'norm': 'trajectory'}): gode_dopri5_traj_fwd /
    gode_dopri5_traj_backprop_bwd.  Same semantics as running torchdiffeq once per trajectory."""

    @staticmethod
    @_nvtx("gode.dopri5_traj.fwd")
    def forward(ctx, y0, meta, W1, b1, W2, b2):
        L = _lib.lib()
        B, D = y0.shape
        H = W1.shape[0]
        T = meta["T"]
        dev = y0.device
        y0c, W1c, b1c, W2c, b2c = (_f32c(x) for x in (y0, W1, b1, W2, b2))
        buf, view = _alloc_traj(T, B, D, meta["layout"], y0)
        o = meta["opts"]
        keep = meta["keep_ckpt"]
        kc = o.ckpt_capacity if keep else 0
        cap = int(meta["traj_log_capacity"])
        opts = GodeAdaptiveOpts.from_buffer_copy(bytes(o))
        opts.ckpt_capacity, opts.log_capacity = kc, cap
        hdr = torch.empty(64, dtype=torch.uint8, device=dev)
        counts = torch.empty((2, B), dtype=torch.int32, device=dev)
        att = torch.empty(13 * cap * B, dtype=torch.uint8, device=dev) if cap > 0 else None
        ckpt = torch.empty((max(kc, 1), B, D), dtype=torch.float32, device=dev) if keep else None
        acc = torch.empty((2, max(kc, 1), B), dtype=torch.float64, device=dev) if keep else None
        ab = att.data_ptr() if att is not None else 0
        tarr = meta["t64"]
        _lib.check(L.gode_dopri5_traj_fwd(
            _ptr(y0c), _ptr(W1c), _ptr(b1c), _ptr(W2c), _ptr(b2c), tarr.ctypes.data, B, D, H, T, C.byref(opts),
            meta["layout"], _ptr(buf), hdr.data_ptr(), counts[0].data_ptr(), counts[1].data_ptr(),
            ab if cap else None, (ab + 8 * cap * B) if cap else None, (ab + 12 * cap * B) if cap else None,
            _ptr(ckpt), acc[0].data_ptr() if keep else None, acc[1].data_ptr() if keep else None, _stream()),
            "gode_dopri5_traj_fwd")
        log = TrajStepLog(hdr, counts[0], counts[1], att, cap, B)
        _LAST_LOG[0] = log
        if meta["check"]:
            raise_for_status(log.status)
        ctx.meta, ctx.log, ctx.kc, ctx.tarr = meta, log, kc, tarr
        ctx.save_for_backward(hdr, counts, ckpt, acc, W1c, b1c, W2c, b2c)
        return view

    @staticmethod
    @_nvtx("gode.dopri5_traj.bwd")
    @_bwd_on_device
    def backward(ctx, grad_traj):
        L = _lib.lib()
        hdr, counts, ckpt, acc, W1c, b1c, W2c, b2c = ctx.saved_tensors
        meta, kc = ctx.meta, ctx.kc
        if ckpt is None:
            raise GodeError("dopri5 forward ran without checkpoints (inputs did not require grad)")
        if meta["check"]:
            raise_for_status(ctx.log.status)
        T = meta["T"]
        _, B, D = ckpt.shape
        H = W1c.shape[0]
        g = _grad_in_layout(grad_traj, meta["layout"])
        grad_y0 = torch.empty((B, D), dtype=torch.float32, device=ckpt.device)
        grad_p = torch.empty(L.gode_param_count(D, H), dtype=torch.float32, device=ckpt.device)
        ws_bytes = L.gode_bwd_workspace_bytes(B, D, H)
        ws = _workspace(ckpt.device, ws_bytes)
        _lib.check(L.gode_dopri5_traj_backprop_bwd(
            _ptr(g), _ptr(W1c), _ptr(b1c), _ptr(W2c), _ptr(b2c), ctx.tarr.ctypes.data, B, D, H, T, meta["layout"],
            hdr.data_ptr(), counts[0].data_ptr(), _ptr(ckpt), acc[0].data_ptr(), acc[1].data_ptr(), kc,
            C.c_float(meta["opts"].fsign), _ptr(grad_y0), _ptr(grad_p), _ptr(ws), ws_bytes, _stream()),
            "gode_dopri5_traj_backprop_bwd")
        needs = ctx.needs_input_grad
        gW1, gb1, gW2, gb2 = _split_params(grad_p, D, H, needs[2:6])
        return (grad_y0 if needs[0] else None), None, gW1, gb1, gW2, gb2


def _adaptive_opts(rtol, atol, options, fsign) -> GodeAdaptiveOpts:
    if isinstance(rtol, torch.Tensor) or isinstance(atol, torch.Tensor):
        rtol, atol = float(rtol), float(atol)
    o = GodeAdaptiveOpts()
    o.rtol, o.atol = float(rtol), float(atol)
    fs = options.get("first_step", None)
    o.first_step = float(fs) if fs is not None else 0.0
    o.safety = float(options.get("safety", 0.9))
    o.ifactor = float(options.get("ifactor", 10.0))
    o.dfactor = float(options.get("dfactor", 0.2))
    o.min_step = float(options.get("min_step", 0.0))
    o.max_step = float(options.get("max_step", float("inf")))
    o.max_num_steps = int(min(options.get("max_num_steps", 2 ** 31 - 1), 2 ** 31 - 1))
    norm = options.get("norm", None)
    if norm is not None and norm not in ("batch", "rms", "trajectory", "per_trajectory", "world"):
        raise NotImplementedError("custom error norms are not supported by the fused dopri5 kernels: 'batch' (torchdiffeq's "
                                  "batch-global RMS, default), 'world' (the same over all data-parallel ranks) or "
                                  "'trajectory' (per-trajectory step control)")
    o.norm_scope = _lib.NORM_TRAJ if norm in ("trajectory", "per_trajectory") else _lib.NORM_BATCH
    o.log_capacity = int(options.get("log_capacity", config.log_capacity))
    o.ckpt_capacity = int(options.get("ckpt_capacity", config.ckpt_capacity))
    o.fsign = fsign
    o.tableau = 0
    return o


# ------------------------------------------------------------------------------------------------------------
# The thin C++ host (csrc_torch/gode_torch.cpp -> _gode_torch.so): autograd node, allocations, workspace and the two C-ABI
# launches of an rk4 / dopri5 call in C++, no Python in the backward.  The Python above turns a call into a PLAN once and
# caches it by everything the plan depends on; per call what is left here is the argument checks, one dict lookup and one
# extension call.  Calls that need something the C++ host does not do (data-parallel exchanges, NVTX ranges, eager status
# checks, euler / midpoint, per-trajectory or world-scope step control, step_size) run on the Python autograd.Functions —
# same C ABI, same kernels.
_ext = [None, False]   # (module, load attempted)
_FRONT = {}            # call key -> (kind, plan id, log capacity) | False (not eligible)


def _load_ext():
    if not _ext[1]:
        _ext[1] = True
        try:
            from . import _gode_torch as m
        except ImportError as e:   # not built: the Python host is complete on its own; say so once
            warnings.warn("gan_ode_b200._gode_torch is not built ({}); using the Python autograd host. "
                          "`python -m gan_ode_b200.build` builds it.".format(e))
            return None
        L = _lib.lib()
        names = ("gode_rk4_fwd", "gode_rk4_adjoint_bwd", "gode_rk4_backprop_bwd", "gode_dopri5_fwd", "gode_dopri5_backprop_bwd",
                 "gode_dopri5_adjoint_bwd", "gode_rk4_bwd_workspace_bytes", "gode_dopri5_workspace_bytes",
                 "gode_dopri5_adjoint_workspace_bytes", "gode_param_count", "gode_stream_capture_id",
                 "gode_set_thread_launch_flags", "gode_strerror")
        m.bind({n: C.cast(getattr(L, n), C.c_void_p).value for n in names})
        mb = _mailbox_view()
        m.set_hooks(GodeError, check_status, _mailbox[0].data_ptr() if mb is not None else 0)
        _ext[0] = m
    return _ext[0]


def _ext_enabled():
    return (config.use_cpp_host and config.grad_allreduce in (None, False) and config.grad_exchange is None
            and not config.nvtx)


def _plan_for(tag, meta, dt, D, H):
    """Turn a prepared call into a C++ plan, or return False if the C++ host does not cover it."""
    m = _load_ext()
    if m is None or meta.get("check") or meta.get("method", 0) or meta.get("world") is not None:
        return False
    if tag == "rk4":
        host = isinstance(dt, np.ndarray)
        pid = m.make_plan(0, meta["T"], meta["layout"], meta["precision"], _bwd_precision(meta, D, H), bool(meta["adjoint"]),
                          dt.tobytes() if host else b"", None if host else dt, b"", b"", b"", 15, False)
        return (0, pid, 0)
    o = meta["opts"]
    if o.tableau:      # bosh3 / adaptive_heun run on the Python Functions
        return False
    if tag == "dopri5":
        pid = m.make_plan(1, meta["T"], meta["layout"], meta["precision"], 0, False, b"", None, meta["t64"].tobytes(), bytes(o),
                          b"", 15, bool(meta["keep_ckpt"]))
        return (1, pid, o.log_capacity)
    if tag == "dopri5_adjoint":
        pid = m.make_plan(2, meta["T"], meta["layout"], meta["precision"], 0, True, b"", None, meta["t64"].tobytes(), bytes(o),
                          bytes(meta["adj_opts"]), meta["param_mask"], False)
        return (2, pid, o.log_capacity, meta["adj_opts"].log_capacity)
    return False


def _run_plan(plan, y0, W1, b1, W2, b2):
    m = _ext[0]
    if plan[0] == 0:
        return m.rk4(y0, W1, b1, W2, b2, plan[1], config.pdl)
    if plan[0] == 1:
        sol, raw = m.dopri5(y0, W1, b1, W2, b2, plan[1], config.pdl)
        _LAST_LOG[0] = StepLog(raw, plan[2])
        return sol
    sol, raw = m.dopri5_adjoint(y0, W1, b1, W2, b2, plan[1])
    _LAST_LOG[0] = StepLog(raw, plan[2])
    _LAST_ADJ_LOG[0] = ("ext", plan[3])
    return sol


def _front_key(y0, t, rtol, atol, method, options, adjoint, adj, weights):
    """Everything a plan depends on, hashable — or None if this call cannot be keyed (device-resident t, tensor tolerances,
    unhashable option values): such calls take the general path."""
    if t.is_cuda or isinstance(rtol, torch.Tensor) or isinstance(atol, torch.Tensor):
        return None
    try:
        okey = tuple(sorted(options.items())) if options else ()
        akey = None
        if adj is not None:
            akey = (adj[0], adj[1], tuple(sorted(adj[2].items())) if adj[2] else adj[2])
        W1 = weights[0]
        grad = torch.is_grad_enabled()
        pmask = sum(1 << k for k, q in enumerate(weights) if q.requires_grad)
        key = (method, adjoint, rtol, atol, okey, akey, t.dtype, t.detach().numpy().tobytes(), W1.shape[1], W1.shape[0],
               grad and (y0.requires_grad or pmask != 0), pmask if grad else 0,
               config.layout, config.precision, config.dopri5_adjoint, config.ckpt_capacity, config.log_capacity)
        hash(key)
        return key
    except TypeError:
        return None


def _solve(func, y0, t, rtol, atol, method, options, adjoint: bool, adj=None):
    W1, b1, W2, b2 = recognise_field(func)
    _check_common(y0, t)
    _require_cuda(y0, weights=(W1, b1, W2, b2))
    if _mailbox_view() is not None:
        check_status()
    with _on_device(y0.device):
        return _solve_on_device(func, y0, t, rtol, atol, method, options, adjoint, adj, (W1, b1, W2, b2))


def _solve_on_device(func, y0, t, rtol, atol, method, options, adjoint, adj, weights):
    W1, b1, W2, b2 = weights
    fkey = None
    if _ext_enabled():
        fkey = _front_key(y0, t, rtol, atol, method, options, adjoint, adj, weights)
        plan = _FRONT.get(fkey) if fkey is not None else None
        if plan:
            if y0.shape[1] != W1.shape[1]:
                raise ValueError("y0 has {} features but func expects {}".format(y0.shape[1], W1.shape[1]))
            return _run_plan(plan, y0, W1, b1, W2, b2)
        if plan is False:
            fkey = None     # known not to be coverable: skip the plan attempt below

    def dispatch(tag, fn, meta, dt=None):
        """The one place a prepared call is launched: through the C++ host if it covers the call, else the Python Function."""
        if fkey is not None:
            plan = _plan_for(tag, meta, dt, D, H)
            if len(_FRONT) < 512:
                _FRONT[fkey] = plan
            if plan:
                return _run_plan(plan, y0, W1, b1, W2, b2)
        if tag == "rk4":
            return fn.apply(y0, dt, meta, W1, b1, W2, b2)
        return fn.apply(y0, meta, W1, b1, W2, b2)

    options = {} if options is None else dict(options)
    if method is None:
        method = "dopri5"
    D, H = W1.shape[1], W1.shape[0]
    if y0.shape[1] != D:
        raise ValueError("y0 has {} features but func expects {}".format(y0.shape[1], D))
    prec_name = options.get("precision", config.precision)
    if prec_name not in _lib.PREC:
        raise ValueError("options['precision'] must be one of {}".format(sorted(_lib.PREC)))
    prec = _lib.PREC[prec_name]
    if (D, H, prec) not in _SUPPORTED:
        if not _lib.lib().gode_supported(D, H, prec) or not _lib.lib().gode_supported(D, H, _lib.PREC["fp32"]):
            raise NotImplementedError("no sm_100a kernel compiled for ODEFunc(dim={}, dim_hidden={}) at precision {}; "
                                      "there is no fallback".format(D, H, prec_name))
        _SUPPORTED.add((D, H, prec))
    layout = _layout_code(options.get("layout", config.layout))
    if t.is_cuda and t.device != y0.device:
        warnings.warn("t is not on the same device as y0. Coercing to y0.device.")
        t = t.to(y0.device)
    meta = dict(T=len(t), layout=layout, precision=prec, adjoint=adjoint, check=bool(options.get("check", False)),
                bwd_precision=options.get("bwd_precision"))

    if method in ("rk4", "euler", "midpoint"):
        if method != "rk4":   # torchdiffeq's other fixed-grid methods on the same field
            if (D, H) != (16, 16) or prec != _lib.PREC["fp32"]:
                raise NotImplementedError('method "{}" exists for the reference shape D=H=16 in fp32 only'.format(method))
            meta["method"] = _lib.METHODS[method]
        step_size = options.get("step_size")
        if step_size is not None:
            if adjoint:
                # torchdiffeq: the adjoint solves inherit the forward options, i.e. every output interval is re-solved on its
                # own step_size grid (gode_fixed_adjoint_bwd_substep); FP32 kernels of the reference shape
                if (D, H) != (16, 16) or prec != _lib.PREC["fp32"]:
                    raise NotImplementedError("options['step_size'] under odeint_adjoint exists for D=H=16 in fp32")
                if adj is not None or t.is_cuda:
                    raise NotImplementedError("step_size under odeint_adjoint: host-resident t, no separate adjoint options")
                m = dict(meta, t=t, step_size=step_size)
                return _FixedSubstepAdjoint.apply(y0, m, W1, b1, W2, b2)
            sub = dict(options)
            sub.pop("step_size")

            def apply_fixed(grid):
                m = dict(meta, T=len(grid), layout=_lib.LAYOUT_TBD)
                return _Rk4.apply(y0, _rk4_dt(grid, sub, y0.device), m, W1, b1, W2, b2)

            return _substepped(apply_fixed, y0, t, step_size, options.get("layout", config.layout))
        dt = _rk4_dt(t, options, y0.device)
        return dispatch("rk4", _Rk4, meta, dt)

    if method in _lib.TABLEAUS:     # dopri5 (torchdiffeq's default) and, SURVEY §8 f4, bosh3 / adaptive_heun on the same kernels
        wide = not (D == 16 and H == 16)   # csrc/wide_dopri5.cu: one warp per trajectory, state in global memory, any batch
        if wide and method != "dopri5":
            raise NotImplementedError('method "{}" exists for the reference shape D=H=16 only (the wide fields have rk4 and '
                                      'dopri5)'.format(method))
        if prec != _lib.PREC["fp32"]:
            raise NotImplementedError("adaptive solvers run in fp32 only: the error estimate is below tf32/bf16 resolution")
        # odeint_adjoint + dopri5 (the ODE-RNN call, models/mocogan_ode_rnn.py:47-48): by default torchdiffeq's continuous
        # adjoint (gode_dopri5_adjoint_bwd).  options={'adjoint': 'discrete'} (or config.dopri5_adjoint) instead
        # differentiates the recorded accepted steps (gode_dopri5_backprop_bwd): one replay, no second adaptive solve, any
        # batch the forward holds; the two gradients agree to O(tolerance).
        # torchdiffeq: t -> float64 for adaptive solvers; the grid rides in the launch parameters (syncs iff t is on GPU)
        if len(t) > _lib.ADAPTIVE_MAX_T:
            raise NotImplementedError("the adaptive kernels take at most {} output times (they travel in the launch "
                                      "parameters); got {} — split t".format(_lib.ADAPTIVE_MAX_T, len(t)))
        t64, _, fsign = _host_steps(t.cpu() if t.is_cuda else t)
        meta["opts"] = _adaptive_opts(rtol, atol, options, fsign)
        meta["opts"].tableau = _lib.TABLEAUS[method]
        meta["t64"] = t64
        meta["traj_log_capacity"] = int(options.get("traj_log_capacity", 0))  # per-attempt logs per trajectory (tests)
        meta["keep_ckpt"] = torch.is_grad_enabled() and (y0.requires_grad or any(p.requires_grad for p in (W1, b1, W2, b2)))
        other = method != "dopri5"
        if wide:
            if meta["opts"].norm_scope == _lib.NORM_TRAJ or options.get("norm") == "world":
                raise NotImplementedError("wide fields: dopri5 with torchdiffeq's batch-global norm only")
            if adjoint and options.get("adjoint", config.dopri5_adjoint) == "continuous" and meta["keep_ckpt"]:
                raise NotImplementedError("wide fields: the continuous dopri5 adjoint re-solve is not built; pass "
                                          "options={'adjoint': 'discrete'} (gradient of the recorded steps) or use odeint")
            return _Dopri5.apply(y0, meta, W1, b1, W2, b2)
        if meta["opts"].norm_scope == _lib.NORM_TRAJ:
            if other:
                raise NotImplementedError("per-trajectory step control exists for dopri5 only")
            if adjoint and meta["keep_ckpt"] and (adj is not None or options.get("adjoint", config.dopri5_adjoint) == "continuous"):
                # (silently handing back the recorded-step gradient would ignore adjoint_rtol / adjoint_atol / adjoint_options)
                raise NotImplementedError("odeint_adjoint with options['norm']='trajectory': the continuous adjoint re-solve uses "
                                          "torchdiffeq's batch-global norm; pass options={'adjoint': 'discrete'} (gradient of "
                                          "the recorded per-trajectory steps, no adjoint_* arguments) or use odeint")
            return dispatch("dopri5_traj", _Dopri5Traj, meta)
        mode = options.get("adjoint", config.dopri5_adjoint)
        if other:
            if options.get("norm") == "world":
                raise NotImplementedError("the world-scope norm exists for dopri5 only")
            if adjoint and mode == "continuous":
                raise NotImplementedError('odeint_adjoint with method="{}": the continuous adjoint re-solve is built for dopri5 '
                                          "only; pass options={{'adjoint': 'discrete'}} (gradient of the recorded steps) or use "
                                          "odeint".format(method))
        if mode not in ("continuous", "discrete"):
            raise ValueError("options['adjoint'] must be 'continuous' or 'discrete'")
        if options.get("norm") == "world":
            if config.world_norm is None:
                raise GodeError("options['norm']='world' needs gan_ode_b200.dist.enable_world_norm() on every rank first")
            if adjoint and mode == "continuous":
                raise NotImplementedError("the world-scope norm exists for the forward solve and its recorded-step gradient: "
                                          "use odeint, or odeint_adjoint with options['adjoint']='discrete'")
            meta["world"] = config.world_norm
        if adjoint and mode == "continuous" and meta["keep_ckpt"]:
            a_rtol, a_atol, a_options = adj if adj is not None else (None, None, None)
            if a_options is None:   # adjoint.py: the forward options without its norm
                a_options = {k: v for k, v in options.items() if k in ("first_step", "safety", "ifactor", "dfactor",
                                                                      "min_step", "max_step", "max_num_steps", "log_capacity")}
            elif "norm" in a_options:
                raise NotImplementedError("a custom adjoint norm is not supported: the kernel uses torchdiffeq's default")
            # adjoint.py: adjoint_params = the parameters with requires_grad; only those are in the augmented state / norm
            meta["param_mask"] = sum(1 << k for k, q in enumerate((W1, b1, W2, b2)) if q.requires_grad)
            meta["adj_opts"] = _adaptive_opts(rtol if a_rtol is None else a_rtol, atol if a_atol is None else a_atol,
                                              dict(a_options), fsign)
            return dispatch("dopri5_adjoint", _Dopri5Adjoint, meta)
        return dispatch("dopri5", _Dopri5, meta)

    raise NotImplementedError('method "{}" is not built (rk4, euler, midpoint, dopri5, bosh3 and adaptive_heun are; '
                              'SURVEY §8f-4)'.format(method))


def odeint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, event_fn=None):
    """torchdiffeq.odeint for the reference's ODEFunc.  Gradients are those of autograd through the solver
    (backprop-through-solver; the adaptive dt sequence is treated as data)."""
    if event_fn is not None:
        raise NotImplementedError("event handling is not on the gan-ode hot path")
    return _solve(func, y0, t, rtol, atol, method, options, adjoint=False)


def odeint_adjoint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, event_fn=None,
                   adjoint_rtol=None, adjoint_atol=None, adjoint_method=None, adjoint_options=None,
                   adjoint_params=None):
    """torchdiffeq.odeint_adjoint for the reference's ODEFunc.  method='rk4': the backward pass is torchdiffeq's
    continuous adjoint re-solved per output interval with the same method (adjoint.py), fused into one kernel.
    method='dopri5' (default): the same continuous adjoint with the adaptive solver (adjoint_rtol / adjoint_atol /
    adjoint_options as upstream), or options={'adjoint': 'discrete'} for the gradient of the recorded steps (see _solve)."""
    if event_fn is not None:
        raise NotImplementedError("event handling is not on the gan-ode hot path")
    if adjoint_params is None and not isinstance(func, nn.Module):
        raise ValueError("func must be an instance of nn.Module to specify the adjoint parameters; alternatively they "
                         "can be specified explicitly via the `adjoint_params` argument. If there are no parameters "
                         "then it is allowable to set `adjoint_params=()`.")
    if adjoint_method is not None and adjoint_method != (method or "dopri5"):
        raise NotImplementedError("adjoint_method must equal method on the fused path")
    is_dopri5 = (method or "dopri5") == "dopri5"
    if not is_dopri5:   # fixed grid: the adjoint re-solve uses the forward grid, tolerances are not used
        if adjoint_options:
            raise NotImplementedError("adjoint_options are not supported with a fixed-grid method on the fused path")
    if adjoint_params is not None:
        mine = field_parameters(func)
        given = [p for p in adjoint_params]
        if len(given) != len(mine) or any(a is not b for a, b in zip(given, mine)):
            raise NotImplementedError("adjoint_params must be func's own parameters (or None)")
    return _solve(func, y0, t, rtol, atol, method, options, adjoint=True,
                  adj=(adjoint_rtol, adjoint_atol, adjoint_options) if is_dopri5 else None)
