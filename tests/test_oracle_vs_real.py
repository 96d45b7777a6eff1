"""Oracle vs the REAL torchdiffeq / torchsde, whenever they are importable (they are not in this image: the whole module
is skipped then, and DESIGN.md §4 says "parity unpinned").  If a later environment provides them — e.g. a wheel dropped
under baseline/_ref — these tests pin the restatement: rk4 bit for bit, dopri5 step for step, adjoint gradients to fp32
round-off, Euler–Maruyama given the same Brownian increments."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REF = os.path.join(ROOT, "baseline", "_ref")
if os.path.isdir(_REF) and _REF not in sys.path:
    sys.path.append(_REF)

real = pytest.importorskip("torchdiffeq", reason="real torchdiffeq not installed (parity unpinned, see DESIGN.md §4)")

from oracle import torchdiffeq_restatement as tdq  # noqa: E402
from tests.helpers import make_field, rel_err  # noqa: E402


def _grads(mod, adjoint, f, y0, t, g, **kw):
    y = y0.clone().requires_grad_(True)
    sol = (mod.odeint_adjoint if adjoint else mod.odeint)(f, y, t, **kw)
    return sol.detach(), torch.autograd.grad((sol * g).sum(), [y] + list(f.parameters()))


@pytest.mark.parametrize("adjoint", [False, True])
def test_rk4_bit_for_bit(adjoint):
    f = make_field(seed=1)
    y0, t, g = torch.randn(32, 16), torch.linspace(0, 1, 16).float(), torch.randn(16, 32, 16)
    s0, g0 = _grads(real, adjoint, f, y0, t, g, method="rk4")
    s1, g1 = _grads(tdq, adjoint, f, y0, t, g, method="rk4")
    assert torch.equal(s0, s1)
    for a, b in zip(g0, g1):
        assert rel_err(b, a) <= 1e-6


@pytest.mark.parametrize("scale", [1.0, 4.0, 8.0])
def test_dopri5_step_for_step(scale):
    f = make_field(seed=2, scale=scale)
    y0, t = torch.randn(64, 16), torch.linspace(0, 1, 16).float()
    nfe = [0]

    class Counted(torch.nn.Module):
        def __init__(self, inner):
            super().__init__()
            self.inner = inner

        def forward(self, tt, yy):
            nfe[0] += 1
            return self.inner(tt, yy)

    with torch.no_grad():
        s0 = real.odeint(Counted(f), y0, t, method="dopri5", rtol=1e-5, atol=1e-5)
        s1 = tdq.odeint(f, y0, t, method="dopri5", rtol=1e-5, atol=1e-5)
    log = tdq.last_step_log()
    assert nfe[0] == log.nfe                      # same number of attempted steps
    assert rel_err(s1, s0) <= 1e-6


def test_dopri5_adjoint_default_tolerances():
    f = make_field(seed=3)
    y0, t, g = torch.randn(16, 16), torch.tensor([0.0, 1.0]), torch.randn(2, 16, 16)
    s0, g0 = _grads(real, True, f, y0, t, g)
    s1, g1 = _grads(tdq, True, f, y0, t, g)
    assert rel_err(s1, s0) <= 1e-6
    for a, b in zip(g0, g1):
        assert rel_err(b, a) <= 1e-5


def test_euler_maruyama_given_increments():
    tsde_real = pytest.importorskip("torchsde", reason="real torchsde not installed")
    from oracle import torchsde_restatement as tsde
    from tests.helpers import SDEFunc
    torch.manual_seed(4)
    sde = SDEFunc(16, 16)
    ts = torch.linspace(0, 1, 16).float()
    y0 = torch.randn(8, 16)
    grid = tsde.step_grid(ts, 2.5e-2)
    dW = torch.randn(len(grid), 8, 16) * torch.tensor([float(b - a) for a, b in grid]).sqrt().view(-1, 1, 1)

    class Table:
        def __init__(self):
            self.k = 0

        def __call__(self, ta, tb):
            w = dW[self.k]
            self.k += 1
            return w

    with torch.no_grad():
        s0 = tsde_real.sdeint(sde, y0, ts, bm=Table(), method="euler", dt=2.5e-2)
        s1 = tsde.sdeint(sde, y0, ts, bm=tsde.TableBrownian(dW), method="euler", dt=2.5e-2)
    assert rel_err(s1, s0) <= 1e-6
