"""CPU tests of host-side logic of the boundary (no GPU, no kernel calls)."""
import numpy as np
import pytest
import torch

import gan_ode_b200 as gode
from gan_ode_b200 import odeint as _unused  # noqa: F401
from gan_ode_b200.sdeint import recognise_sde, step_grid
from oracle import torchsde_restatement as tsde
from oracle.latent_motion import ODEFunc, SDEFunc
import importlib

api = importlib.import_module("gan_ode_b200.odeint")


def test_sde_step_grid_equals_torchsde_restatement():
    for T, dt in ((16, 2.5e-2), (16, 0.1), (5, 0.3), (2, 1e-2)):
        ts = torch.linspace(0, 1, T).float()
        h, out_step, w0, w1 = step_grid(ts, dt)
        pairs = tsde.step_grid(ts, dt)
        assert len(h) == len(pairs)
        assert np.array_equal(h, np.array([float(b - a) for a, b in pairs], dtype=np.float32))
        assert out_step[0] == 0 and (np.diff(out_step[1:]) >= 0).all()
        assert np.allclose(w0[1:] + w1[1:], 1.0, atol=1e-6)
    h, *_ = step_grid(torch.linspace(0, 1, 16).float(), 2.5e-2)
    assert len(h) == 41


def test_host_steps_matches_torch_arithmetic():
    t = torch.linspace(0, 1, 16).float()
    t64, dt, fsign = api._host_steps(t)
    assert fsign == 1.0 and np.array_equal(dt, (t[1:] - t[:-1]).numpy())
    assert len(set(dt.tolist())) > 1  # fp32 linspace spacing is not uniform
    t64, dt, fsign = api._host_steps(torch.linspace(1, 0, 9))
    assert fsign == -1.0 and (dt < 0).all() and (np.diff(t64) > 0).all()
    t64, dt, _ = api._host_steps(torch.linspace(0, 1, 16, dtype=torch.float64))
    assert dt.dtype == np.float32
    with pytest.raises(AssertionError):
        api._host_steps(torch.tensor([0.0, 0.5, 0.5, 1.0]))


def test_field_recognition():
    W1, b1, W2, b2 = gode.recognise_field(ODEFunc(16, 24))
    assert W1.shape == (24, 16) and W2.shape == (16, 24)

    class Wrong(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.fn = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.ReLU(), torch.nn.Linear(4, 4))

    with pytest.raises(NotImplementedError):
        gode.recognise_field(Wrong())

    class OneLayer(torch.nn.Module):   # models/mocogan_mnist.py:6-16
        def __init__(self):
            super().__init__()
            self.fn = torch.nn.Sequential(torch.nn.Linear(16, 16), torch.nn.Tanh())

        def forward(self, t, x):
            return self.fn(x)

    one = OneLayer()
    W1, b1, W2, b2 = gode.recognise_field(one)
    assert W1 is one.fn[0].weight and torch.equal(W2, torch.eye(16)) and not W2.requires_grad and not b2.any()
    from gan_ode_b200.odeint import field_parameters
    assert [id(q) for q in field_parameters(one)] == [id(q) for q in one.parameters()]
    f, g = recognise_sde(SDEFunc(16, 16))
    assert f[0].shape == (16, 16) and g[2].shape == (16, 16)
    with pytest.raises(NotImplementedError):
        recognise_sde(ODEFunc(16, 16))


def test_field_recognition_checks_the_forward_not_only_the_structure():
    """A func that HAS the reference's fn stack but whose forward is something else (time-dependent, negated, residual) must
    not be integrated as fn(x); forward hooks would be skipped by the kernels, so they are refused too."""
    def variant(body):
        class F(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.fn = torch.nn.Sequential(torch.nn.Linear(16, 16), torch.nn.Tanh(), torch.nn.Linear(16, 16))
            forward = body
        return F()

    for body in (lambda self, t, x: -self.fn(x), lambda self, t, x: self.fn(x) * torch.cos(t), lambda self, t, x: x + self.fn(x),
                 lambda self, t, x: self.fn(x)[:, :8]):
        with pytest.raises(NotImplementedError, match="does not return"):
            gode.recognise_field(variant(body))
    good = variant(lambda self, t, x: self.fn(x))
    gode.recognise_field(good)
    h = good.fn[1].register_forward_hook(lambda m, i, o: o * 2)
    with pytest.raises(NotImplementedError, match="hooks"):
        gode.recognise_field(good)
    h.remove()
    gode.recognise_field(good)

    class TimeSDE(SDEFunc):
        def g(self, t, x):
            return self.diffusion_fn(x) * (1 + t)

    with pytest.raises(NotImplementedError, match="does not return"):
        recognise_sde(TimeSDE(16, 16))


def test_table_limits_are_reported_with_their_real_names(monkeypatch):
    """The step / frame / output-time tables travel in the launch parameters; exceeding one is reported by the host with the
    actual numbers (torchsde's default dt = 1e-3 on [0, 1] is 1000 steps), not by a generic error from the C side."""
    from gan_ode_b200 import _lib
    sdeint_mod = importlib.import_module("gan_ode_b200.sdeint")
    api = importlib.import_module("gan_ode_b200.odeint")
    monkeypatch.setattr(api, "_require_cuda", lambda *a, **k: None)
    monkeypatch.setattr(_lib.lib(), "gode_supported", lambda *a: 1, raising=False)
    sde, y0 = SDEFunc(16, 16), torch.randn(4, 16)
    with pytest.raises(NotImplementedError, match="100[01] Euler-Maruyama steps.*at most 320 steps and 64 frames"):
        gode.sdeint(sde, y0, torch.linspace(0, 1, 16))                      # torchsde's default dt
    with pytest.raises(NotImplementedError, match="at most 320 steps and 64 frames"):
        gode.sdeint(sde, y0, torch.linspace(0, 1, 100), dt=0.05)
    with pytest.raises(NotImplementedError, match="at most 256 output times"):
        gode.odeint(ODEFunc(16, 16), torch.randn(4, 16), torch.linspace(0, 1, 300), method="dopri5")
    assert b"GODE_SDE_MAX" in _lib.lib().gode_strerror(-3) and b"GODE_ADAPTIVE_MAX_T" in _lib.lib().gode_strerror(-3)


def test_boundary_rejects_cpu_and_bad_inputs_without_touching_the_gpu():
    f = ODEFunc(16, 16)
    t = torch.linspace(0, 1, 16)
    with pytest.raises(gode.GodeError):
        gode.odeint(f, torch.randn(4, 16), t, method="rk4")
    with pytest.raises(TypeError):
        gode.odeint(f, torch.zeros(4, 16, dtype=torch.long), t)
    with pytest.raises(NotImplementedError):
        gode.odeint(f, torch.randn(4, 16), t, method="rk4", event_fn=lambda t, y: y)
    with pytest.raises(gode.GodeError):
        gode.sdeint(SDEFunc(16, 16), torch.randn(4, 16), t, method="euler", dt=0.025)


def test_shims_register_both_packages():
    import sys
    saved = {k: sys.modules.pop(k, None) for k in ("torchdiffeq", "torchsde")}
    try:
        gode.install_shims()
        from torchdiffeq import odeint_adjoint
        from torchsde import sdeint_adjoint
        assert odeint_adjoint is gode.odeint_adjoint and sdeint_adjoint is gode.sdeint_adjoint
    finally:
        for k, v in saved.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v


def test_grad_exchange_hook_prefers_the_fused_kernel_for_small_vectors_only():
    """odeint._maybe_allreduce: a callable in config.grad_allreduce (gan_ode_b200.dist.P2PAllReduce on a GPU box) gets the
    small flat gradient buffers; anything larger than its `small` threshold goes to torch.distributed."""
    import importlib
    import torch
    api = importlib.import_module("gan_ode_b200.odeint")

    calls = []

    class Fake:
        small, cap = 1024, 65536

        def __call__(self, flat):
            calls.append(flat.numel())
            return flat

    prev = api.config.grad_allreduce
    api.config.grad_allreduce = Fake()
    try:
        api._maybe_allreduce(torch.zeros(544))
        assert calls == [544]
        import pytest
        with pytest.raises(Exception):  # no process group here: the large vector must have been routed to dist.all_reduce
            api._maybe_allreduce(torch.zeros(4096))
        assert calls == [544]
    finally:
        api.config.grad_allreduce = prev


def test_nvtx_wrapper_is_transparent_to_autograd():
    """config.nvtx wraps the autograd forward/backward of the boundary in NVTX ranges; with the flag off (default) the
    wrapper must be invisible to torch.autograd.Function.apply, needs_input_grad and the gradient flow."""
    import importlib
    api = importlib.import_module("gan_ode_b200.odeint")

    class Scale(torch.autograd.Function):
        @staticmethod
        @api._nvtx("test.fwd")
        def forward(ctx, x, k):
            ctx.k = k
            return x * k

        @staticmethod
        @api._nvtx("test.bwd")
        def backward(ctx, g):
            assert ctx.needs_input_grad == (True, False)
            return g * ctx.k, None

    assert api.config.nvtx is False
    x = torch.randn(5, requires_grad=True)
    y = Scale.apply(x, 3.0)
    y.sum().backward()
    assert torch.equal(x.grad, torch.full((5,), 3.0))
    for cls in (api._Rk4, api._Dopri5, api._Dopri5Adjoint, api._Dopri5Traj):
        assert cls.forward.__name__ == "forward" and cls.backward.__name__ == "backward"


def test_sde_adjoint_grid_matches_the_oracle_grids():
    """gan_ode_b200.sdeint.adjoint_grid (host tables of gode_sde_em_fwd_cells / gode_sde_adjoint_bwd) against the oracle's
    independently written grids: same cell times, same reverse steps, cells tile every step exactly."""
    import numpy as np
    from gan_ode_b200.sdeint import adjoint_grid
    from oracle import torchsde_restatement as tsde
    for ts, dt in ((torch.linspace(0, 1, 16).float(), 2.5e-2), (torch.tensor([0.0, 0.3, 0.35, 1.1]), 0.07)):
        ag = adjoint_grid(ts, dt)
        og = tsde.adjoint_time_grid(ts, dt).numpy()
        assert ag.R == len(og) - 1 and np.abs(ag.times - og).max() == 0.0
        fwd = tsde.step_grid(ts, dt)
        assert len(fwd) == len(ag.fwd[0]) == len(ag.fwd_lo) - 1
        for k, (a, b) in enumerate(fwd):
            assert ag.times[ag.fwd_lo[k]] == float(a) and ag.times[ag.fwd_lo[k + 1]] == float(b)
        rev = [(i, s0, s1) for i, pairs in tsde.reverse_step_grid(ts, dt) for s0, s1 in pairs]
        assert len(rev) == ag.n_rev
        for n, (i, s0, s1) in enumerate(rev):
            assert ag.ibeg[i] <= n < ag.iend[i]
            assert ag.times[ag.rev_hi[n]] == float(-s0) and ag.times[ag.rev_lo[n]] == float(-s1)
            assert ag.h_rev[n] == np.float32(float(s1 - s0))
        assert np.allclose(ag.cell_sqrt ** 2, np.diff(og), rtol=1e-6, atol=0)
