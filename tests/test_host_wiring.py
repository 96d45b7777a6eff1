"""Host-side wiring of the boundary, on CPU: the Python host (torchdiffeq signatures -> autograd Functions -> ctypes) is run
end to end with the COMPUTE entry points of libgode.so replaced by recorders (the real library still answers every size /
capability query).  Checks what crosses the C ABI — entry point, sizes, layout and precision codes, which tensors' pointers —
and what comes back to autograd (shapes, strides, which gradients), without a GPU and without computing anything.  The
numerical parity of the kernels themselves is the -m gpu suite."""
import ctypes
import importlib

import pytest
import torch

import gan_ode_b200 as gode
from gan_ode_b200 import _lib
from tests.helpers import make_field

api = importlib.import_module("gan_ode_b200.odeint")

COMPUTE = ("gode_rk4_fwd", "gode_rk4_adjoint_bwd", "gode_rk4_backprop_bwd", "gode_fixed_fwd", "gode_fixed_adjoint_bwd",
           "gode_fixed_backprop_bwd", "gode_dopri5_fwd", "gode_dopri5_backprop_bwd", "gode_dopri5_adjoint_bwd",
           "gode_dopri5_traj_fwd", "gode_dopri5_traj_backprop_bwd", "gode_sde_em_fwd", "gode_sde_em_bwd", "gode_sde_em_fwd_cells",
           "gode_sde_adjoint_bwd", "gode_rk4_sampler_fwd", "gode_rk4_adjoint_bwd_strided", "gode_fixed_adjoint_bwd_substep", "gode_odernn_fwd",
           "gode_odernn_bwd", "gode_gru_jump_fwd", "gode_gru_jump_bwd")


class Recorder:
    """libgode.so with the compute entry points swapped for call recorders that return 0 (GODE_OK)."""

    def __init__(self):
        self.real = _lib.lib()
        self.calls = []

    def __getattr__(self, name):
        if name in COMPUTE:
            def rec(*args):
                self.calls.append((name, args))
                return 0
            return rec
        return getattr(self.real, name)


@pytest.fixture()
def wired(monkeypatch):
    r = Recorder()
    monkeypatch.setattr(_lib, "lib", lambda: r)
    monkeypatch.setattr(api, "_stream", lambda: 0)
    monkeypatch.setattr(importlib.import_module("gan_ode_b200.odernn"), "_stream", lambda: 0)   # (imported by name there)

    monkeypatch.setattr(api, "_require_cuda", lambda *a, **k: None)   # the device gate is the one thing switched off here
    # these tests watch the PYTHON host cross the C ABI; the C++ host (gan_ode_b200._gode_torch) binds the real entry points
    # by address and needs CUDA tensors — it is what the -m gpu suite runs by default
    monkeypatch.setattr(api.config, "use_cpp_host", False)
    return r


def _ints(args):
    return [a for a in args if isinstance(a, int) and not isinstance(a, bool)]


def test_rk4_adjoint_call_crosses_the_abi_as_documented(wired):
    f = make_field(seed=1)
    y0 = torch.randn(6, 16, requires_grad=True)
    t = torch.linspace(0, 1, 16)
    sol = gode.odeint_adjoint(f, y0, t, method="rk4")
    assert sol.shape == (16, 6, 16) and sol.is_contiguous()
    (name, a), = wired.calls
    assert name == "gode_rk4_fwd"
    W1, b1, W2, b2 = gode.recognise_field(f)
    assert list(a[:5]) == [y0.data_ptr(), W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr()]
    assert list(a[6:13]) == [0, 6, 16, 16, 16, _lib.PREC["fp32"], _lib.LAYOUT_TBD]     # dt on host | B D H T | precision layout
    dt = (ctypes.c_float * 15).from_address(a[5])
    assert list(dt) == (t[1:] - t[:-1]).tolist()                                      # step table by value, fp32
    assert a[13] == sol.data_ptr()
    grads = torch.autograd.grad(sol.sum(), [y0] + list(f.parameters()))
    assert [g.shape for g in grads] == [(6, 16), (16, 16), (16,), (16, 16), (16,)]
    name, a = wired.calls[1]
    assert name == "gode_rk4_adjoint_bwd" and a[0] == sol.data_ptr()                   # continuous adjoint reads the stored outputs
    # plain odeint: backprop-through-solver entry point; no gradient requested for y0 -> None comes back for it
    wired.calls.clear()
    sol = gode.odeint(f, y0.detach(), t, method="rk4")
    torch.autograd.grad(sol.sum(), list(f.parameters()))
    assert [c[0] for c in wired.calls] == ["gode_rk4_fwd", "gode_rk4_backprop_bwd"]


def test_layout_precision_and_method_options_reach_the_abi(wired):
    f = make_field(seed=2)
    y0 = torch.randn(5, 16)
    t = torch.tensor([1.0, 0.6, 0.0])                                                 # decreasing grid
    sol = gode.odeint(f, y0, t, method="rk4", options={"layout": "btd", "precision": "bf16"})
    name, a = wired.calls[-1]
    assert sol.shape == (3, 5, 16) and sol.stride() == (16, 48, 1)                    # (T,B,D) view of a (B,T,D) buffer
    assert list(a[7:13]) == [5, 16, 16, 3, _lib.PREC["bf16"], _lib.LAYOUT_BTD]
    dt = (ctypes.c_float * 2).from_address(a[5])
    assert list(dt) == pytest.approx([-0.4, -0.6])                                    # signed steps for a decreasing grid
    gode.odeint(f, y0, t, method="midpoint")
    name, a = wired.calls[-1]
    assert name == "gode_fixed_fwd" and a[0] == _lib.METHODS["midpoint"]
    with pytest.raises(NotImplementedError):
        gode.odeint(f, y0, t, method="dopri8")
    # step_size under the adjoint: fine-grid forward launch, then the sub-stepped adjoint entry point with its device tables
    wired.calls.clear()
    yr = y0.clone().requires_grad_(True)
    sol = gode.odeint_adjoint(f, yr, t, method="rk4", options={"step_size": 0.1})
    assert [c[0] for c in wired.calls] == ["gode_rk4_fwd"] and wired.calls[0][1][10] == 11     # 1.0 -> 0.0 in steps of 0.1: 11 points
    sol.sum().backward()
    name, a = wired.calls[-1]
    assert name == "gode_fixed_adjoint_bwd_substep" and a[0] == _lib.METHODS["rk4"] and list(a[10:15]) == [5, 16, 16, 3, _lib.LAYOUT_TBD]
    sub_dt, beg, end = api._adjoint_substeps(t, 0.1, y0.device)
    assert beg.tolist() == [0, 6, 0] and end.tolist() == [0, 10, 6]      # interval [0.0, 0.6]: 6 steps, then [0.6, 1.0]: 4
    assert (sub_dt < 0).all() and abs(float(sub_dt[:6].sum()) + 0.6) < 1e-6                   # decreasing t: negative steps



def test_dopri5_calls_continuous_adjoint_discrete_gradient_and_per_trajectory(wired):
    f = make_field(seed=3)
    t = torch.tensor([0.0, 1.0])

    def names():
        out = [c[0] for c in wired.calls]
        wired.calls.clear()
        return out

    # the ODE-RNN call: odeint_adjoint with torchdiffeq's defaults -> forward without checkpoints + continuous adjoint
    y0 = torch.randn(4, 16, requires_grad=True)
    sol = gode.odeint_adjoint(f, y0, t)
    name, a = wired.calls[0]
    opts = a[10]._obj
    assert (opts.rtol, opts.atol, opts.ckpt_capacity, opts.norm_scope) == (1e-7, 1e-9, 0, _lib.NORM_BATCH)
    assert a[18] is None and a[19] is None                                            # no checkpoint buffers
    torch.autograd.grad(sol.sum(), [y0] + list(f.parameters()))
    name, a = wired.calls[1]
    assert name == "gode_dopri5_adjoint_bwd" and a[0] == sol.data_ptr() and a[13] == 15  # all four tensors in the norm
    assert (a[12]._obj.rtol, a[12]._obj.atol) == (1e-7, 1e-9)
    names()
    # adjoint tolerances of their own; a frozen parameter leaves the augmented state (param_mask)
    f.fn[2].bias.requires_grad_(False)
    sol = gode.odeint_adjoint(f, y0, t, rtol=1e-5, atol=1e-6, adjoint_rtol=1e-3, adjoint_atol=1e-4,
                              adjoint_options={"first_step": 0.05})
    torch.autograd.grad(sol.sum(), [y0])
    name, a = wired.calls[1]
    assert (a[12]._obj.rtol, a[12]._obj.atol, a[12]._obj.first_step, a[13]) == (1e-3, 1e-4, 0.05, 7)
    f.fn[2].bias.requires_grad_(True)
    names()
    # opt-in gradient of the recorded steps, and odeint (backprop through the solver): checkpoints + replay kernel
    for call in (lambda: gode.odeint_adjoint(f, y0, t, options={"adjoint": "discrete"}), lambda: gode.odeint(f, y0, t)):
        sol = call()
        assert wired.calls[0][1][10]._obj.ckpt_capacity == api.config.ckpt_capacity
        torch.autograd.grad(sol.sum(), [y0])
        assert names() == ["gode_dopri5_fwd", "gode_dopri5_backprop_bwd"]
    # no gradient wanted: no checkpoints are kept
    with torch.no_grad():
        gode.odeint(f, y0, t)
    assert wired.calls[0][1][10]._obj.ckpt_capacity == 0
    names()
    # per-trajectory step control
    sol = gode.odeint(f, y0, t, options={"norm": "trajectory"})
    torch.autograd.grad(sol.sum(), [y0])
    assert names() == ["gode_dopri5_traj_fwd", "gode_dopri5_traj_backprop_bwd"]
    with pytest.raises(gode.GodeError, match="enable_world_norm"):
        gode.odeint(f, y0, t, options={"norm": "world"})
    with pytest.raises(ValueError):
        gode.odeint_adjoint(f, y0, t, options={"adjoint": "exact"})


def test_single_layer_field_and_nvtx_flag_do_not_change_what_is_called(wired, monkeypatch):
    class OneLayer(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.fn = torch.nn.Sequential(torch.nn.Linear(16, 16), torch.nn.Tanh())

        def forward(self, t, x):
            return self.fn(x)

    f = OneLayer()
    y0 = torch.randn(3, 16, requires_grad=True)
    t = torch.linspace(0, 1, 4)
    pushed = []
    monkeypatch.setattr(torch.cuda.nvtx, "range_push", lambda s: pushed.append(s))
    monkeypatch.setattr(torch.cuda.nvtx, "range_pop", lambda: pushed.append("pop"))
    monkeypatch.setattr(api.config, "nvtx", True)
    sol = gode.odeint_adjoint(f, y0, t, method="rk4")
    grads = torch.autograd.grad(sol.sum(), [y0] + list(f.parameters()))
    assert len(grads) == 3 and grads[1].shape == (16, 16) and grads[2].shape == (16,)   # only the two real parameters
    name, a = wired.calls[0]
    W2 = torch.eye(16)
    got = torch.frombuffer((ctypes.c_float * 256).from_address(a[3]), dtype=torch.float32).view(16, 16)
    assert torch.equal(got, W2)                                                       # identity second layer goes to the kernel
    assert pushed == ["gode.rk4.fwd", "pop", "gode.rk4.bwd", "pop"]


def test_sde_call_carries_the_torchsde_step_grid_and_the_philox_contract(wired):
    from gan_ode_b200.sdeint import PhiloxBrownian, TableBrownian
    from tests.helpers import SDEFunc
    torch.manual_seed(0)
    sde = SDEFunc(16, 16)
    y0 = torch.randn(7, 16, requires_grad=True)
    ts = torch.linspace(0, 1, 16)
    # the reference's call (models/mocogan_sde.py:57-59): forward on the cell grid, backward = torchsde's stochastic adjoint
    sol = gode.sdeint_adjoint(sde, y0, ts, bm=PhiloxBrownian(1234, traj_offset=70), method="euler", adjoint_method="euler", dt=2.5e-2)
    assert sol.shape == (16, 7, 16)
    name, a = wired.calls[0]
    assert name == "gode_sde_em_fwd_cells" and a[0] == y0.data_ptr() and a[4] == 41
    assert a[10] == 79 and list(a[11:15]) == [7, 16, 16, 16] and a[15] is None and (a[16], a[17]) == (1234, 70)   # R | B D H T | Philox
    grads = torch.autograd.grad(sol.sum(), [y0] + list(sde.parameters()))
    name, a = wired.calls[1]
    assert len(grads) == 9 and name == "gode_sde_adjoint_bwd"
    assert a[4] == 45 and a[11] == 79 and list(a[12:16]) == [7, 16, 16, 16] and (a[17], a[18]) == (1234, 70)       # 45 reverse steps
    wired.calls.clear()
    # options={'adjoint': 'discrete'}: one draw per forward step, exact gradient of the recorded steps
    sol = gode.sdeint_adjoint(sde, y0, ts, bm=PhiloxBrownian(1234, traj_offset=70), method="euler", adjoint_method="euler", dt=2.5e-2,
                              options={"adjoint": "discrete"})
    name, a = wired.calls[0]
    assert name == "gode_sde_em_fwd" and a[0] == y0.data_ptr()
    assert a[4] == 41                                              # torchsde's fp32 time accumulation: 41 steps, not 40
    assert list(a[8:12]) == [7, 16, 16, 16] and a[12] is None      # B D H T | no increment table: in-kernel Philox
    assert (a[13], a[14]) == (1234, 70)                            # seed, global index of this shard's first trajectory
    drift = list(a[1])
    assert drift[0] == sde.drift_fn[0].weight.data_ptr() and list(a[2])[0] == sde.diffusion_fn[0].weight.data_ptr()
    grads = torch.autograd.grad(sol.sum(), [y0] + list(sde.parameters()))
    assert len(grads) == 9 and wired.calls[1][0] == "gode_sde_em_bwd" and wired.calls[1][1][14:16] == (1234, 70)
    # given increments: the table's pointer crosses instead
    dW = torch.randn(41, 7, 16)
    gode.sdeint(sde, y0, ts, bm=TableBrownian(dW), method="euler", dt=2.5e-2)
    assert wired.calls[-1][1][12] == dW.data_ptr()
    with pytest.raises(ValueError):
        gode.sdeint(sde, y0, ts, bm=TableBrownian(dW[:40]), method="euler", dt=2.5e-2)
    with pytest.raises(NotImplementedError):
        gode.sdeint(sde, y0, ts, method="milstein", dt=2.5e-2)


def test_odernn_sampler_calls_one_entry_point_per_direction(wired):
    f = make_field(seed=4)
    cell = torch.nn.GRUCell(16, 16)
    h0 = torch.randn(9, 16, requires_grad=True)
    eps = torch.randn(5, 9, 16)
    codes = gode.odernn_codes(f, cell, h0, eps)
    assert codes.shape == (5, 9, 16)
    name, a = wired.calls[0]
    assert name == "gode_odernn_fwd" and list(a[10:14]) == [9, 16, 16, 5]           # B D H F
    o = a[14]._obj
    assert (o.rtol, o.atol, o.norm_scope) == (1e-7, 1e-9, _lib.NORM_BATCH) and o.ckpt_capacity > 0
    assert a[15] == codes.data_ptr() and a[20] is None                             # no per-trajectory counts in batch mode
    torch.autograd.grad(codes.sum(), [h0] + list(f.parameters()) + list(cell.parameters()))
    name, a = wired.calls[1]
    assert name == "gode_odernn_bwd" and a[20] is None and a[21] is None            # recorded-step gradient: no adjoint options
    wired.calls.clear()
    # torchdiffeq's continuous adjoint per frame: no checkpoints in the forward, adjoint tolerances to the backward
    codes = gode.odernn_codes(f, cell, h0, eps, rtol=1e-6, atol=1e-8, options={"adjoint": "continuous"})
    assert wired.calls[0][1][14]._obj.ckpt_capacity == 0
    torch.autograd.grad(codes.sum(), [h0])
    a = wired.calls[1][1]
    assert (a[21]._obj.rtol, a[21]._obj.atol) == (1e-6, 1e-8)
    assert a[22] == 15                                                               # all four tensors are adjoint parameters
    wired.calls.clear()
    # adjoint_params = the parameters that require grad (adjoint.py): frozen tensors leave the augmented state and its norm
    f.fn[0].bias.requires_grad_(False)
    f.fn[2].weight.requires_grad_(False)
    codes = gode.odernn_codes(f, cell, h0, eps, options={"adjoint": "continuous"})
    torch.autograd.grad(codes.sum(), [h0])
    assert wired.calls[1][1][22] == 0b1001                                           # W1 and b2 only
    f.fn[0].bias.requires_grad_(True)
    f.fn[2].weight.requires_grad_(True)
    wired.calls.clear()
    # per-trajectory step control: the counts buffer crosses
    codes = gode.odernn_codes(f, cell, h0, eps, options={"norm": "trajectory"})
    assert wired.calls[0][1][14]._obj.norm_scope == _lib.NORM_TRAJ and wired.calls[0][1][20] is not None
    with pytest.raises(NotImplementedError):
        gode.odernn_codes(f, torch.nn.GRUCell(16, 16, bias=False), h0, eps)
    out = gode.gru_jump(eps[0], h0, cell)
    assert out.shape == (9, 16) and wired.calls[-1][0] == "gode_gru_jump_fwd"


def test_device_gate_rejects_cpu_field_and_cpu_state():
    """the gate covers the weights too.  Without a GPU in this container the CUDA-tensor half is emulated with
    the `meta`-free check itself: a CPU y0 raises; a y0 that passes as CUDA with CPU weights must raise before any pointer
    crosses the C ABI."""
    f = make_field(seed=2)
    with pytest.raises(gode.GodeError, match="no CPU fallback"):
        gode.odeint(f, torch.randn(4, 16), torch.linspace(0, 1, 4), method="rk4")

    class FakeCuda:
        """Stands in for a CUDA tensor in the gate (is_cuda / device only)."""
        is_cuda = True
        device = torch.device("cuda", 0)

    with pytest.raises(gode.GodeError, match="parameters are on cpu"):
        api._require_cuda(FakeCuda(), weights=tuple(f.parameters()))
    other = FakeCuda()
    other.device = torch.device("cuda", 1)
    with pytest.raises(gode.GodeError, match="different devices"):
        api._require_cuda(FakeCuda(), other, what="h0 / eps")


def test_fused_sampler_writes_into_the_cat_buffer_and_reads_gradients_in_place(wired):
    """one forward launch for randn + pre-MLP + solve + transpose + cat; the adjoint reads the codes and the
    upstream gradient inside z / grad_z through their row stride."""
    from tests.caller_model import LatentMotionODE
    torch.manual_seed(0)
    m = LatentMotionODE(16, 16)
    content = torch.randn(5, 50)
    z = gode.fused_sample_z(m.linear, m.ode_fn, 5, 16, content=content, seed=77, traj_offset=1000)
    assert z.shape == (5 * 16, 66) and torch.equal(z.view(5, 16, 66)[:, 3, :50], content)
    name, a = wired.calls[-1]
    assert name == "gode_rk4_sampler_fwd"
    assert a[0] == m.linear[0].weight.data_ptr() and a[2] == m.linear[2].weight.data_ptr() and abs(a[4] - 0.2) < 1e-7 and a[5] == 64
    assert a[6] == m.ode_fn.fn[0].weight.data_ptr() and list(a[12:16]) == [5, 16, 16, 16]
    assert (a[16], a[17], a[18]) == (77, 1000, None) and a[19] == _lib.LAYOUT_BTD
    assert a[20] == z.data_ptr() + 4 * 50 and a[21] == 66                     # motion columns of z, row stride 66 floats
    w = torch.randn_like(z)
    grads = torch.autograd.grad((z * w).sum(), list(m.linear.parameters()) + list(m.ode_fn.parameters()))
    assert [g.shape for g in grads] == [(64, 16), (64,), (16, 64), (16,), (16, 16), (16,), (16, 16), (16,)]
    name, a = wired.calls[-1]
    assert name == "gode_rk4_adjoint_bwd_strided" and a[0] == z.data_ptr() + 200 and a[1] == 66 and a[3] == 66
    # subset of a larger batch: global trajectory ids ride along; nn.Identity pre-MLP -> no prologue weights
    ids = torch.tensor([3, 900, 17])
    codes = gode.fused_sample_z(torch.nn.Identity(), m.ode_fn, 3, 16, seed=5, traj_ids=ids)
    name, a = wired.calls[-1]
    assert codes.shape == (48, 16) and a[0] is None and a[5] == 0 and a[18] is not None and a[21] == 16
    with pytest.raises(NotImplementedError):
        gode.fused_sample_z(torch.nn.Linear(16, 16), m.ode_fn, 3, 16)
