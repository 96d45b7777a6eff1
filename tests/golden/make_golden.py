"""Generate the golden input/output vectors under tests/golden/*.npz.

    python tests/golden/make_golden.py            # in the build container (needs /root/reference for the caller cases)

The reference's solver dependency (torchdiffeq / torchsde) is not installable here, so the numbers are
produced by the CPU oracle (oracle/torchdiffeq_restatement.py, oracle/torchsde_restatement.py) under fixed seeds.  The
"caller" cases go through the UNMODIFIED reference modules (/root/reference/models/mocogan_ode.py, mocogan_ode_rnn.py)
with the oracle registered as `torchdiffeq`, so they freeze the reference's own wiring (pre-MLP, time grid, frame-major
reshape, ODE-RNN loop) together with the solver arithmetic.  Every case stores the inputs it was made from, so the
CUDA path is checked on the GPU box (where neither /root/reference nor this generator's environment exists) without
re-deriving anything.  tests/test_golden.py holds both checks: oracle vs golden (CPU) and kernels vs golden (GPU).
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import torchdiffeq_restatement as tdq  # noqa: E402
from oracle import torchsde_restatement as tsde  # noqa: E402
from oracle.latent_motion import ODEFunc, SDEFunc  # noqa: E402

REF = "/root/reference"
PNAMES = ("W1", "b1", "W2", "b2")


def _np(x):
    return x.detach().cpu().numpy().copy()


def _field(D, H, seed, scale=1.0):
    torch.manual_seed(seed)
    f = ODEFunc(D, H)
    if scale != 1.0:
        with torch.no_grad():
            for p in f.parameters():
                p.mul_(scale)
    return f


def _params(f, prefix=""):
    return {prefix + n: _np(p) for n, p in zip(PNAMES, f.parameters())}


def _solve_case(solver, f, y0, t, g, **kw):
    y = y0.clone().requires_grad_(True)
    sol = solver(f, y, t, **kw)
    grads = torch.autograd.grad((sol * g).sum(), [y] + list(f.parameters()))
    out = dict(y0=_np(y0), t=_np(t), grad_traj=_np(g), sol=_np(sol), grad_y0=_np(grads[0]))
    out.update(_params(f))
    out.update({"grad_" + n: _np(x) for n, x in zip(PNAMES, grads[1:])})
    return out


def case_rk4(adjoint, B=8, D=16, H=16, seed=11, t=None):
    f = _field(D, H, seed)
    torch.manual_seed(seed + 1)
    y0 = torch.randn(B, D)
    t = torch.linspace(0, 1, 16).float() if t is None else t
    g = torch.randn(len(t), B, D)
    return _solve_case(tdq.odeint_adjoint if adjoint else tdq.odeint, f, y0, t, g, method="rk4")


def case_dopri5_backprop(B=8, seed=21, scale=4.0):
    """configs[1]'s call at a small batch: dopri5 rtol=atol=1e-5, gradients of autograd through the solver with the
    step sizes treated as data (oracle switch _detach_dt0)."""
    f = _field(16, 16, seed, scale)
    torch.manual_seed(seed + 1)
    y0 = torch.randn(B, 16)
    t = torch.linspace(0, 1, 16).float()
    g = torch.randn(16, B, 16)
    out = _solve_case(tdq.odeint, f, y0, t, g, method="dopri5", rtol=1e-5, atol=1e-5, options={"_detach_dt0": True})
    log = tdq.last_step_log()
    out.update(accepted=np.array(log.accepted, dtype=np.uint8), dt=np.array(log.dt, dtype=np.float64),
               error_ratio=np.array(log.error_ratio, dtype=np.float64), nfe=np.array(log.nfe))
    return out


def case_dopri5_adjoint_default_tol(B=4, seed=31):
    """The ODE-RNN call (models/mocogan_ode_rnn.py:47-48): odeint_adjoint(func, h, [0,1]) with torchdiffeq defaults."""
    f = _field(16, 16, seed)
    torch.manual_seed(seed + 1)
    y0 = torch.randn(B, 16)
    t = torch.tensor([0.0, 1.0])
    g = torch.randn(2, B, 16)
    out = _solve_case(tdq.odeint_adjoint, f, y0, t, g)
    log = tdq.last_step_log()
    return out


def case_sde(B=8, seed=41):
    torch.manual_seed(seed)
    sde = SDEFunc(16, 16)
    ts = torch.linspace(0, 1, 16).float()
    grid = tsde.step_grid(ts, 2.5e-2)
    h = torch.tensor([float(b - a) for a, b in grid])
    torch.manual_seed(seed + 1)
    y0 = torch.randn(B, 16)
    g = torch.randn(16, B, 16)
    dW = torch.randn(len(grid), B, 16) * h.sqrt().view(-1, 1, 1)
    y = y0.clone().requires_grad_(True)
    sol = tsde.sdeint(sde, y, ts, bm=tsde.TableBrownian(dW), method="euler", dt=2.5e-2)
    names = [n for n, _ in sde.named_parameters()]
    grads = torch.autograd.grad((sol * g).sum(), [y] + list(sde.parameters()))
    out = dict(y0=_np(y0), t=_np(ts), grad_traj=_np(g), dW=_np(dW), sol=_np(sol), grad_y0=_np(grads[0]),
               param_names=np.array(names))
    for n, p, gp in zip(names, sde.parameters(), grads[1:]):
        out["p:" + n] = _np(p)
        out["g:" + n] = _np(gp)
    return out


def case_sde_stochastic_adjoint(B=8, seed=43):
    """models/mocogan_sde.py:57-59 with the Brownian path given on the cell grid: forward + torchsde's stochastic adjoint."""
    torch.manual_seed(seed)
    sde = SDEFunc(16, 16)
    ts = torch.linspace(0, 1, 16).float()
    times = tsde.adjoint_time_grid(ts, 2.5e-2)
    torch.manual_seed(seed + 1)
    y0 = torch.randn(B, 16)
    g = torch.randn(16, B, 16)
    dW = torch.randn(len(times) - 1, B, 16) * torch.diff(times).float().sqrt().view(-1, 1, 1)
    y = y0.clone().requires_grad_(True)
    sol = tsde.sdeint_adjoint(sde, y, ts, bm=tsde.GridBrownian(times, dW), method="euler", adjoint_method="euler", dt=2.5e-2)
    names = [n for n, _ in sde.named_parameters()]
    grads = torch.autograd.grad((sol * g).sum(), [y] + list(sde.parameters()))
    out = dict(y0=_np(y0), t=_np(ts), grad_traj=_np(g), times=times.numpy(), dW=_np(dW), sol=_np(sol), grad_y0=_np(grads[0]),
               param_names=np.array(names))
    for n, p, gp in zip(names, sde.parameters(), grads[1:]):
        out["p:" + n] = _np(p)
        out["g:" + n] = _np(gp)
    return out


def _reference_modules():
    """Import the unmodified reference model files with the oracle as `torchdiffeq` (and `on_dev` aliased to `models`,
    SURVEY Appendix C)."""
    shim = types.ModuleType("torchdiffeq")
    shim.odeint, shim.odeint_adjoint = tdq.odeint, tdq.odeint_adjoint
    sys.modules["torchdiffeq"] = shim
    sys.path.insert(0, REF)
    torch.cuda.is_available = lambda: False
    ode = importlib.import_module("models.mocogan_ode")
    import models
    sys.modules["on_dev"] = models
    sys.modules["on_dev.mocogan_ode"] = ode
    rnn = importlib.import_module("models.mocogan_ode_rnn")
    return ode, rnn


def case_reference_sample_z_m(ode_mod, n=6, seed=51):
    """VideoGeneratorMNISTODE.sample_z_m (models/mocogan_ode.py:133-148) run from the reference file."""
    torch.manual_seed(seed)
    gen = ode_mod.VideoGeneratorMNISTODE(1, 50, 0, 16, 16)
    torch.manual_seed(seed + 1)
    noise = torch.randn(n, 16)          # what sample_z_m draws first (models/mocogan_ode.py:136)
    torch.manual_seed(seed + 1)
    codes = gen.sample_z_m(n)
    torch.manual_seed(seed + 2)
    g = torch.randn_like(codes)
    names = ["ode_fn.fn.0.weight", "ode_fn.fn.0.bias", "ode_fn.fn.2.weight", "ode_fn.fn.2.bias",
             "linear.0.weight", "linear.0.bias", "linear.2.weight", "linear.2.bias"]
    sd = dict(gen.named_parameters())
    grads = torch.autograd.grad((codes * g).sum(), [sd[k] for k in names])
    out = dict(noise=_np(noise), codes=_np(codes), grad_codes=_np(g), param_names=np.array(names))
    for k, gp in zip(names, grads):
        out["p:" + k] = _np(sd[k])
        out["g:" + k] = _np(gp)
    return out


def case_reference_odernn(rnn_mod, n=3, T=4, seed=61):
    """VideoGeneratorMNISTODERNN.sample_z_m (models/mocogan_ode_rnn.py:40-54) run from the reference file; the noise it
    draws (h0 then one e_t per frame, models/mocogan.py:297-301) is re-derived from the same seed and stored."""
    torch.manual_seed(seed)
    gen = rnn_mod.VideoGeneratorMNISTODERNN(1, 50, 0, 16, T)
    torch.manual_seed(seed + 1)
    h0 = torch.randn(n, 16)
    eps = torch.stack([torch.randn(n, 16) for _ in range(T)])
    torch.manual_seed(seed + 1)
    codes = gen.sample_z_m(n)
    torch.manual_seed(seed + 2)
    g = torch.randn_like(codes)
    names = ["ode_fn.fn.0.weight", "ode_fn.fn.0.bias", "ode_fn.fn.2.weight", "ode_fn.fn.2.bias",
             "recurrent.weight_ih", "recurrent.weight_hh", "recurrent.bias_ih", "recurrent.bias_hh"]
    sd = dict(gen.named_parameters())
    grads = torch.autograd.grad((codes * g).sum(), [sd[k] for k in names])
    out = dict(h0=_np(h0), eps=_np(eps), codes=_np(codes), grad_codes=_np(g), param_names=np.array(names))
    for k, gp in zip(names, grads):
        out["p:" + k] = _np(sd[k])
        out["g:" + k] = _np(gp)
    return out


def main():
    cases = {
        "rk4_adjoint_B8": case_rk4(True),
        "rk4_backprop_B8": case_rk4(False, seed=12),
        "rk4_adjoint_nonuniform_decreasing": case_rk4(True, B=5, seed=13,
                                                       t=torch.tensor([1.0, 0.8, 0.75, 0.4, 0.1, 0.0])),
        "rk4_adjoint_wide_D64_H256_B4": case_rk4(True, B=4, D=64, H=256, seed=14),
        "dopri5_backprop_tol1e-5_B8": case_dopri5_backprop(),
        "dopri5_adjoint_default_tol_B4": case_dopri5_adjoint_default_tol(),
        "sde_euler_given_dW_B8": case_sde(),
        "sde_stochastic_adjoint_B8": case_sde_stochastic_adjoint(),
    }
    if os.path.isdir(os.path.join(REF, "models")):
        ode_mod, rnn_mod = _reference_modules()
        cases["reference_sample_z_m_ode"] = case_reference_sample_z_m(ode_mod)
        cases["reference_sample_z_m_odernn"] = case_reference_odernn(rnn_mod)
    else:
        print("WARNING: /root/reference absent - caller cases not regenerated")
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    for name, arrays in cases.items():
        if only and name not in only:
            continue
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **arrays)
        print("{:45s} {:8.1f} KB".format(name, os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
