"""Host-side wiring of the thin C++ host (gan_ode_b200/_gode_torch.so, csrc_torch/gode_torch.cpp) on CPU.

The C++ host reaches libgode.so only through entry-point ADDRESSES handed over by `bind`.  Here the compute entry points are
bound to ctypes callbacks that record their arguments and return a chosen code, so the C++ autograd nodes run end to end on
CPU tensors: what crosses the C ABI (entry point, sizes, layout / precision codes, pointers, option structs), what comes back
to autograd, that it matches the Python host call for call, and that C-ABI failures surface as the same Python exception
types — including from a backward.  Numerical parity of the kernels behind it is the -m gpu suite."""
import ctypes as C
import importlib

import pytest
import torch

import gan_ode_b200 as gode
from gan_ode_b200 import _lib
from tests.helpers import make_field

api = importlib.import_module("gan_ode_b200.odeint")
STUBBED = ("gode_rk4_fwd", "gode_rk4_adjoint_bwd", "gode_rk4_backprop_bwd", "gode_dopri5_fwd", "gode_dopri5_backprop_bwd",
           "gode_dopri5_adjoint_bwd")
ALL = STUBBED + ("gode_rk4_bwd_workspace_bytes", "gode_dopri5_workspace_bytes", "gode_dopri5_adjoint_workspace_bytes",
                 "gode_param_count", "gode_stream_capture_id", "gode_set_thread_launch_flags", "gode_strerror")


@pytest.fixture()
def cpp(monkeypatch):
    m = api._load_ext()
    assert m is not None, "gan_ode_b200/_gode_torch.so is not built (python -m gan_ode_b200.build)"
    L = _lib.lib()
    real = {n: C.cast(getattr(L, n), C.c_void_p).value for n in ALL}
    calls, keep, rets = [], [], {}

    def stub(name):
        res, args = _lib._SIGS[name]
        proto = C.CFUNCTYPE(res, *args)

        def fn(*a):
            # option structs live on the caller's stack: snapshot them while the call is in progress
            a = tuple(type(x.contents).from_buffer_copy(bytes(x.contents)) if hasattr(x, "contents") and x else x for x in a)
            calls.append((name, a))
            return rets.get(name, 0)
        cb = proto(fn)
        keep.append(cb)
        return C.cast(cb, C.c_void_p).value

    m.bind(dict(real, **{n: stub(n) for n in STUBBED}))
    m.set_hooks(gode.GodeError, api.check_status, 0)
    monkeypatch.setattr(api, "_require_cuda", lambda *a, **k: None)
    monkeypatch.setattr(api, "_FRONT", {})
    yield calls, rets
    m.bind(real)
    mb = api._mailbox_view()
    m.set_hooks(gode.GodeError, api.check_status, api._mailbox[0].data_ptr() if mb is not None else 0)


def test_rk4_calls_cross_the_abi_from_cpp_exactly_as_from_python(cpp):
    calls, rets = cpp
    f = make_field(seed=1)
    W1, b1, W2, b2 = gode.recognise_field(f)
    y0 = torch.randn(6, 16, requires_grad=True)
    t = torch.linspace(0, 1, 16)
    sol = gode.odeint_adjoint(f, y0, t, method="rk4")
    assert sol.shape == (16, 6, 16) and sol.is_contiguous()
    (name, a), = calls
    assert name == "gode_rk4_fwd"
    assert list(a[:5]) == [y0.data_ptr(), W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr()]
    assert list(a[6:13]) == [0, 6, 16, 16, 16, _lib.PREC["fp32"], _lib.LAYOUT_TBD] and a[13] == sol.data_ptr()
    assert list((C.c_float * 15).from_address(a[5])) == (t[1:] - t[:-1]).tolist()      # the plan's host step table, by value
    grads = torch.autograd.grad(sol.sum(), [y0] + list(f.parameters()))
    assert [g.shape for g in grads] == [(6, 16), (16, 16), (16,), (16, 16), (16,)]
    name, a = calls[1]
    assert name == "gode_rk4_adjoint_bwd" and a[0] == sol.data_ptr() and list(a[8:14]) == [6, 16, 16, 16, 0, 0]
    assert a[14] == grads[0].data_ptr() and a[17] >= _lib.SYNC_REGION_BYTES            # grad_y0 written in place; workspace size
    # the second identical call is served from the plan cache: same plan, no new plan entry
    n_plans = len(api._FRONT)
    gode.odeint_adjoint(f, y0, t, method="rk4")
    assert len(api._FRONT) == n_plans == 1 and all(api._FRONT.values())
    # odeint: the backprop entry point; (B,T,D) layout and a decreasing grid; only the requested gradients come back
    calls.clear()
    f.fn[0].bias.requires_grad_(False)
    sol = gode.odeint(f, y0.detach(), torch.tensor([1.0, 0.6, 0.0]), method="rk4", options={"layout": "btd", "precision": "bf16"})
    assert sol.shape == (3, 6, 16) and sol.stride() == (16, 48, 1)
    assert list(calls[0][1][7:13]) == [6, 16, 16, 3, _lib.PREC["bf16"], _lib.LAYOUT_BTD]
    sol.sum().backward()
    assert calls[1][0] == "gode_rk4_backprop_bwd" and f.fn[0].bias.grad is None and f.fn[0].weight.grad is not None


def test_dopri5_plans_options_and_logs(cpp):
    calls, rets = cpp
    f = make_field(seed=3)
    y0 = torch.randn(4, 16, requires_grad=True)
    t = torch.tensor([0.0, 1.0])
    # the ODE-RNN call (models/mocogan_ode_rnn.py:47-48): torchdiffeq defaults, continuous adjoint, no checkpoints kept
    sol = gode.odeint_adjoint(f, y0, t, adjoint_rtol=1e-3, adjoint_atol=1e-4, adjoint_options={"first_step": 0.05})
    name, a = calls[0]
    opts = a[10]
    assert name == "gode_dopri5_fwd" and (opts.rtol, opts.atol, opts.ckpt_capacity) == (1e-7, 1e-9, 0)
    assert a[18] is None and a[19] is None and a[22] >= _lib.SYNC_REGION_BYTES
    torch.autograd.grad(sol.sum(), [y0] + list(f.parameters()))
    name, a = calls[1]
    aopts = a[12]
    assert name == "gode_dopri5_adjoint_bwd" and a[0] == sol.data_ptr() and a[13] == 15
    assert (aopts.rtol, aopts.atol, aopts.first_step) == (1e-3, 1e-4, 0.05)
    # odeint: checkpoints + replay kernel; the step log object points at the tensor the C++ host allocated
    calls.clear()
    sol = gode.odeint(f, y0, torch.linspace(0, 1, 16), method="dopri5", rtol=1e-5, atol=1e-5)
    name, a = calls[0]
    assert a[10].ckpt_capacity == api.config.ckpt_capacity and a[18] is not None
    assert gode.last_step_log()._raw.data_ptr() == a[13]
    torch.autograd.grad(sol.sum(), [y0])
    assert calls[1][0] == "gode_dopri5_backprop_bwd" and calls[1][1][15] == api.config.ckpt_capacity
    # what the C++ host does not cover stays on the Python Functions (here: they would hit the REAL library, so only the
    # decision is checked)
    for kw in (dict(method="euler"), dict(method="dopri5", options={"norm": "trajectory"}), dict(method="dopri5", options={"check": True})):
        key = api._front_key(y0, t, 1e-7, 1e-9, kw.get("method"), kw.get("options"), False, None, gode.recognise_field(f))
        assert key not in api._FRONT


def test_cabi_failures_raise_the_python_hosts_exception_types_even_from_a_backward(cpp):
    calls, rets = cpp
    f = make_field(seed=5)
    y0 = torch.randn(3, 16, requires_grad=True)
    rets["gode_rk4_backprop_bwd"] = -1          # GODE_ERR_SHAPE
    sol = gode.odeint(f, y0, torch.linspace(0, 1, 4), method="rk4")
    with pytest.raises(gode.GodeError, match="no kernel compiled"):
        sol.sum().backward()
    rets["gode_dopri5_adjoint_bwd"] = _lib.ERR_COOP
    sol = gode.odeint_adjoint(f, y0, torch.tensor([0.0, 1.0]))
    with pytest.raises(gode.GodeError, match="discrete"):
        sol.sum().backward()
    rets["gode_rk4_fwd"] = -4                   # GODE_ERR_WORKSPACE
    with pytest.raises(gode.GodeError, match="workspace too small"):
        gode.odeint_adjoint(f, y0, torch.linspace(0, 1, 4), method="rk4")
