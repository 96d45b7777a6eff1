"""Developer probe: one dopri5 forward + continuous-adjoint backward (two-point grid, default tolerances) at batch argv[1];
run under ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gan_ode_b200 as gode
from tests.helpers import make_field, clone_to

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
f = clone_to(make_field(seed=1), "cuda")
t = torch.tensor([0.0, 1.0])
y = torch.randn(B, 16, device="cuda", requires_grad=True)
for _ in range(2):
    sol = gode.odeint_adjoint(f, y, t)
    torch.autograd.grad(sol.sum(), [y] + list(f.parameters()))
torch.cuda.synchronize()
print("ok", gode.last_adjoint_log().n_attempts)
