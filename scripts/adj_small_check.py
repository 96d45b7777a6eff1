"""Developer probe: tensor-core adjoint (D=H=16, bf16) vs the FP32 adjoint kernel on the same trajectory, + timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gan_ode_b200 as gode
from tests.helpers import make_field, clone_to, rel_err

B = int(sys.argv[1]) if len(sys.argv) > 1 else 300
f = clone_to(make_field(16, 16, seed=1), "cuda")
t = torch.linspace(0, 1, 16).float()
torch.manual_seed(0)
y0 = torch.randn(B, 16, device="cuda")
g = torch.randn(16, B, 16, device="cuda")


def grads(bwd):
    y = y0.clone().requires_grad_(True)
    sol = gode.odeint_adjoint(f, y, t, method="rk4", options={"precision": "bf16", "bwd_precision": bwd})
    return torch.autograd.grad((sol * g).sum(), [y] + list(f.parameters()))


ref = grads("fp32")
out = grads("bf16")
torch.cuda.synchronize()
for n, a, b in zip(["y0", "W1", "b1", "W2", "b2"], out, ref):
    print("%-3s rel err %.3e  (max |ref| %.3e)" % (n, rel_err(a, b), float(b.abs().max())))
for bwd in ("fp32", "bf16"):
    y = y0.clone().requires_grad_(True)
    sol = gode.odeint_adjoint(f, y, t, method="rk4", options={"precision": "bf16", "bwd_precision": bwd})
    ts = []
    for i in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); torch.autograd.grad(sol, [y] + list(f.parameters()), g, retain_graph=True); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    print("B=%d adjoint bwd %s: %.1f us  (%.1f TFLOP/s of 48DH)" % (B, bwd, ms * 1e3, B * 15 * 12288 / ms * 1e-9))
