"""Golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py in the build container).

Two checks per case:
  * CPU (`-m "not gpu"`): the oracle, run again on the stored inputs, reproduces the stored outputs — so an edit to the
    oracle that changes its numbers is caught, and the fixtures stay tied to the committed generator;
  * GPU (`-m gpu`): the CUDA path on the stored inputs equals the stored outputs within BASELINE.json's tolerance
    (1e-5 relative in FP32 mode; accepted-step sequence identical for dopri5).  Nothing here reads /root/reference.
The "reference_*" cases were produced by the unmodified reference model files driving the oracle; the tests replay
them through tests/caller_model.py (the stand-in that test_reference_dropin_cpu.py proves equal to the real files).
"""
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle import torchdiffeq_restatement as tdq
from oracle import torchsde_restatement as tsde
from oracle.latent_motion import ODEFunc, SDEFunc
from tests.helpers import rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PN = ("W1", "b1", "W2", "b2")
ODE_CASES = {
    "rk4_adjoint_B8": dict(adjoint=True, kw=dict(method="rk4")),
    "rk4_backprop_B8": dict(adjoint=False, kw=dict(method="rk4")),
    "rk4_adjoint_nonuniform_decreasing": dict(adjoint=True, kw=dict(method="rk4")),
    "rk4_adjoint_wide_D64_H256_B4": dict(adjoint=True, kw=dict(method="rk4")),
    "dopri5_backprop_tol1e-5_B8": dict(adjoint=False, kw=dict(method="dopri5", rtol=1e-5, atol=1e-5)),
    "dopri5_adjoint_default_tol_B4": dict(adjoint=True, kw=dict()),
}


def load(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False))


def T(x, dev="cpu"):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def field_from(z, dev="cpu"):
    H, D = z["W1"].shape
    f = ODEFunc(D, H)
    with torch.no_grad():
        for p, n in zip(f.parameters(), PN):
            p.copy_(T(z[n]))
    return f.to(dev)


def run_ode(mod, z, adjoint, kw, dev="cpu", extra_options=None):
    f = field_from(z, dev)
    y = T(z["y0"], dev).requires_grad_(True)
    kw = dict(kw)
    if extra_options:
        kw["options"] = dict(kw.get("options", {}), **extra_options)
    sol = (mod.odeint_adjoint if adjoint else mod.odeint)(f, y, T(z["t"]), **kw)
    grads = torch.autograd.grad((sol * T(z["grad_traj"], dev)).sum(), [y] + list(f.parameters()))
    return sol.detach(), grads


def test_every_fixture_is_present_and_small():
    names = sorted(f for f in os.listdir(GOLD) if f.endswith(".npz"))
    assert len(names) == 10, names
    assert sum(os.path.getsize(os.path.join(GOLD, f)) for f in names) < 1 << 20


@pytest.mark.parametrize("name", sorted(ODE_CASES))
def test_oracle_reproduces_golden_ode(name):
    z, c = load(name), ODE_CASES[name]
    opts = {"_detach_dt0": True} if name.startswith("dopri5_backprop") else None
    sol, grads = run_ode(tdq, z, c["adjoint"], c["kw"], extra_options=opts)
    assert rel_err(sol, T(z["sol"])) <= 1e-6
    assert rel_err(grads[0], T(z["grad_y0"])) <= 1e-5
    for g, n in zip(grads[1:], PN):
        assert rel_err(g, T(z["grad_" + n])) <= 1e-5, n
    if "accepted" in z:
        log = tdq.last_step_log()
        assert [int(a) for a in log.accepted] == z["accepted"].tolist()
        assert int(log.nfe) == int(z["nfe"])


def _sde_from(z, dev="cpu"):
    sde = SDEFunc(16, 16)
    with torch.no_grad():
        for n, p in sde.named_parameters():
            p.copy_(T(z["p:" + n]))
    return sde.to(dev)


def test_oracle_reproduces_golden_sde():
    z = load("sde_euler_given_dW_B8")
    sde = _sde_from(z)
    y = T(z["y0"]).requires_grad_(True)
    sol = tsde.sdeint(sde, y, T(z["t"]), bm=tsde.TableBrownian(T(z["dW"])), method="euler", dt=2.5e-2)
    grads = torch.autograd.grad((sol * T(z["grad_traj"])).sum(), [y] + list(sde.parameters()))
    assert z["dW"].shape[0] == 41
    assert rel_err(sol, T(z["sol"])) <= 1e-6
    assert rel_err(grads[0], T(z["grad_y0"])) <= 1e-5
    for (n, _), g in zip(sde.named_parameters(), grads[1:]):
        assert rel_err(g, T(z["g:" + n])) <= 1e-5, n


def test_oracle_reproduces_golden_sde_stochastic_adjoint():
    z = load("sde_stochastic_adjoint_B8")
    sde = _sde_from(z)
    y = T(z["y0"]).requires_grad_(True)
    bm = tsde.GridBrownian(torch.from_numpy(z["times"]), T(z["dW"]))
    sol = tsde.sdeint_adjoint(sde, y, T(z["t"]), bm=bm, method="euler", adjoint_method="euler", dt=2.5e-2)
    grads = torch.autograd.grad((sol * T(z["grad_traj"])).sum(), [y] + list(sde.parameters()))
    assert z["dW"].shape[0] == 79
    assert rel_err(sol, T(z["sol"])) <= 1e-6
    assert rel_err(grads[0], T(z["grad_y0"])) <= 1e-5
    for (n, _), g in zip(sde.named_parameters(), grads[1:]):
        assert rel_err(g, T(z["g:" + n])) <= 1e-5, n


@pytest.mark.gpu
def test_cuda_sde_stochastic_adjoint_matches_golden():
    _need_gpu()
    import gan_ode_b200 as gode
    z = load("sde_stochastic_adjoint_B8")
    sde = _sde_from(z, DEV)
    y = T(z["y0"], DEV).requires_grad_(True)
    sol = gode.sdeint_adjoint(sde, y, T(z["t"]), bm=gode.GridBrownian(z["times"], T(z["dW"], DEV)), method="euler",
                              adjoint_method="euler", dt=2.5e-2)
    grads = torch.autograd.grad((sol * T(z["grad_traj"], DEV)).sum(), [y] + list(sde.parameters()))
    assert rel_err(sol, T(z["sol"])) <= 1e-5
    assert rel_err(grads[0], T(z["grad_y0"])) <= 2e-5
    for (n, _), g in zip(sde.named_parameters(), grads[1:]):
        assert rel_err(g, T(z["g:" + n])) <= 2e-5, n


def _caller_ode(z, dev="cpu"):
    from tests.caller_model import LatentMotionODE
    m = LatentMotionODE(16, 16)
    m.load_state_dict({str(k): T(z["p:" + str(k)]) for k in z["param_names"]})
    return m.to(dev)


def _caller_rnn(z, dev="cpu"):
    from tests.caller_model import LatentMotionODERNN
    m = LatentMotionODERNN(16, z["eps"].shape[0])
    m.load_state_dict({str(k): T(z["p:" + str(k)]) for k in z["param_names"]})
    return m.to(dev)


def _check_caller(z, model, codes, tol_codes, tol_grads, dev="cpu"):
    (codes * T(z["grad_codes"], dev)).sum().backward()
    assert rel_err(codes, T(z["codes"])) <= tol_codes
    sd = dict(model.named_parameters())
    for k in z["param_names"]:
        assert rel_err(sd[str(k)].grad, T(z["g:" + str(k)])) <= tol_grads, (str(k), rel_err(sd[str(k)].grad, T(z["g:" + str(k)])))


@pytest.fixture()
def oracle_as_torchdiffeq(monkeypatch):
    shim = types.ModuleType("torchdiffeq")
    shim.odeint, shim.odeint_adjoint = tdq.odeint, tdq.odeint_adjoint
    monkeypatch.setitem(sys.modules, "torchdiffeq", shim)


def test_standin_caller_on_oracle_reproduces_reference_run_ode(oracle_as_torchdiffeq):
    z = load("reference_sample_z_m_ode")
    m = _caller_ode(z)
    _check_caller(z, m, m.sample_z_m(z["noise"].shape[0], noise=T(z["noise"])), 1e-6, 1e-5)


def test_standin_caller_on_oracle_reproduces_reference_run_odernn(oracle_as_torchdiffeq):
    z = load("reference_sample_z_m_odernn")
    m = _caller_rnn(z)
    _check_caller(z, m, m.sample_z_m(z["h0"].shape[0], h0=T(z["h0"]), eps=T(z["eps"])), 1e-6, 1e-4)


# ---- GPU: the CUDA path against the same fixtures -------------------------------------------------------------------------
DEV = "cuda"


def _need_gpu():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device (no fallback)"


@pytest.mark.gpu
@pytest.mark.parametrize("name", [n for n in sorted(ODE_CASES) if n.startswith("rk4")])
def test_cuda_rk4_matches_golden(name):
    _need_gpu()
    import gan_ode_b200 as gode
    z, c = load(name), ODE_CASES[name]
    sol, grads = run_ode(gode, z, c["adjoint"], c["kw"], dev=DEV)
    assert torch.equal(sol[0].cpu(), T(z["y0"]))
    assert rel_err(sol, T(z["sol"])) <= 1e-5
    assert rel_err(grads[0], T(z["grad_y0"])) <= 1e-5
    for g, n in zip(grads[1:], PN):
        assert rel_err(g, T(z["grad_" + n])) <= 1e-5, n


@pytest.mark.gpu
def test_cuda_dopri5_backprop_matches_golden():
    _need_gpu()
    import gan_ode_b200 as gode
    z = load("dopri5_backprop_tol1e-5_B8")
    sol, grads = run_ode(gode, z, False, dict(method="dopri5", rtol=1e-5, atol=1e-5), dev=DEV)
    log = gode.last_step_log()
    assert [int(a) for a in log.accepted] == z["accepted"].tolist()      # identical accept/reject sequence
    assert int(log.nfe) == int(z["nfe"])
    # trajectory error within the solver tolerance (the two dt sequences agree to fp32 noise in error_ratio^-1/5)
    assert rel_err(sol, T(z["sol"])) <= 1e-5
    assert rel_err(grads[0], T(z["grad_y0"])) <= 1e-4
    for g, n in zip(grads[1:], PN):
        assert rel_err(g, T(z["grad_" + n])) <= 1e-4, n


@pytest.mark.gpu
@pytest.mark.parametrize("mode,tol", [("continuous", 5e-5), ("discrete", 1e-3)])
def test_cuda_dopri5_adjoint_default_tolerances_matches_golden(mode, tol):
    """The ODE-RNN call: forward within fp32 noise of the oracle; gradients by the continuous adjoint kernel (the algorithm
    of the fixture) and, opt-in, by the discrete adjoint of the recorded steps (equal to O(tolerance))."""
    _need_gpu()
    import gan_ode_b200 as gode
    z = load("dopri5_adjoint_default_tol_B4")
    sol, grads = run_ode(gode, z, True, dict(options={"adjoint": mode}), dev=DEV)
    assert rel_err(sol, T(z["sol"])) <= 1e-5
    assert rel_err(grads[0], T(z["grad_y0"])) <= tol
    for g, n in zip(grads[1:], PN):
        assert rel_err(g, T(z["grad_" + n])) <= tol, n


@pytest.mark.gpu
def test_cuda_sde_matches_golden():
    _need_gpu()
    import gan_ode_b200 as gode
    z = load("sde_euler_given_dW_B8")
    sde = _sde_from(z, DEV)
    y = T(z["y0"], DEV).requires_grad_(True)
    sol = gode.sdeint_adjoint(sde, y, T(z["t"]), bm=gode.TableBrownian(T(z["dW"], DEV)), method="euler",
                              adjoint_method="euler", dt=2.5e-2, options={"adjoint": "discrete"})
    grads = torch.autograd.grad((sol * T(z["grad_traj"], DEV)).sum(), [y] + list(sde.parameters()))
    assert rel_err(sol, T(z["sol"])) <= 1e-5
    assert rel_err(grads[0], T(z["grad_y0"])) <= 2e-5
    for (n, _), g in zip(sde.named_parameters(), grads[1:]):
        assert rel_err(g, T(z["g:" + n])) <= 2e-5, n


@pytest.fixture()
def cuda_as_torchdiffeq():
    import gan_ode_b200 as gode
    gode.install_shims()
    yield
    sys.modules.pop("torchdiffeq", None)
    sys.modules.pop("torchsde", None)


@pytest.mark.gpu
def test_cuda_caller_reproduces_reference_run_ode(cuda_as_torchdiffeq):
    _need_gpu()
    z = load("reference_sample_z_m_ode")
    m = _caller_ode(z, DEV)
    _check_caller(z, m, m.sample_z_m(z["noise"].shape[0], noise=T(z["noise"])), 1e-5, 2e-5, DEV)


@pytest.mark.gpu
def test_cuda_caller_reproduces_reference_run_odernn(cuda_as_torchdiffeq):
    _need_gpu()
    z = load("reference_sample_z_m_odernn")
    m = _caller_rnn(z, DEV)
    # through the shim the backward is the continuous dopri5 adjoint, as in the run that made the fixture
    _check_caller(z, m, m.sample_z_m(z["h0"].shape[0], h0=T(z["h0"]), eps=T(z["eps"])), 2e-5, 1e-4, DEV)


@pytest.mark.gpu
def test_cuda_fused_odernn_reproduces_reference_run():
    """The one-call sampler (gode.odernn_codes: C-level frame loop + fused GRU jump) on the fixture made by the
    unmodified models/mocogan_ode_rnn.py."""
    _need_gpu()
    import gan_ode_b200 as gode
    z = load("reference_sample_z_m_odernn")
    m = _caller_rnn(z, DEV)
    codes = gode.odernn_codes(m.ode_fn, m.recurrent, T(z["h0"], DEV), T(z["eps"], DEV))
    _check_caller(z, m, codes.transpose(0, 1).reshape(-1, 16), 2e-5, 1e-3, DEV)
