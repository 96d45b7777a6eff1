"""Developer probe: how much of the eager fwd+bwd call at B=32 is PyTorch's own autograd.Function machinery (floor) and how
much is this package's Python."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gan_ode_b200 as gode
from tests.helpers import make_field, clone_to

f = clone_to(make_field(16, 16, seed=0), "cuda")
t = torch.linspace(0, 1, 16).float()
params = list(f.parameters())
y0 = torch.randn(32, 16, device="cuda", requires_grad=True)
g = torch.randn(16, 32, 16, device="cuda")


class Floor(torch.autograd.Function):  # same arity as _Rk4: 7 inputs, 1 output; two empty() per direction, no kernels
    @staticmethod
    def forward(ctx, y, dt, meta, W1, b1, W2, b2):
        out = torch.empty((16, 32, 16), device=y.device)
        ctx.save_for_backward(out, W1, b1, W2, b2)
        return out

    @staticmethod
    def backward(ctx, go):
        out, W1, b1, W2, b2 = ctx.saved_tensors
        gy = torch.empty((32, 16), device=go.device)
        gp = torch.empty(544, device=go.device)
        return gy, None, None, gp[:256].view(16, 16), gp[256:272], gp[272:528].view(16, 16), gp[528:]


def timeit(fn, n=500):
    for _ in range(30):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


def ours():
    sol = gode.odeint_adjoint(f, y0, t, method="rk4")
    return torch.autograd.grad(sol, [y0] + params, g)


def floor():
    sol = Floor.apply(y0, None, None, *params)
    return torch.autograd.grad(sol, [y0] + params, g)


def fwd_only():
    with torch.no_grad():
        return gode.odeint_adjoint(f, y0, t, method="rk4")


print("fwd+bwd ours %.1f us | autograd.Function floor %.1f us | forward only (no_grad) %.1f us" % (timeit(ours), timeit(floor), timeit(fwd_only)))
