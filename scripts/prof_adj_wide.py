"""Developer probe: one tensor-core adjoint backward of the wide field (D=64, H=256, bf16) at batch argv[1]; run under ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gan_ode_b200 as gode
from tests.helpers import make_field, clone_to

B = int(sys.argv[1]) if len(sys.argv) > 1 else 18944
f = clone_to(make_field(64, 256, seed=0), "cuda")
t = torch.linspace(0, 1, 16).float()
y = torch.randn(B, 64, device="cuda", requires_grad=True)
g = torch.randn(16, B, 64, device="cuda")
sol = gode.odeint_adjoint(f, y, t, method="rk4", options={"precision": "bf16"})
for _ in range(4):
    torch.autograd.grad(sol, [y] + list(f.parameters()), g, retain_graph=True)
torch.cuda.synchronize()
print("ok")
