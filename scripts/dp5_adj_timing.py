import os, sys
sys.path.insert(0, '/root/repo')
import torch
import gan_ode_b200 as gode
from tests.helpers import make_field, clone_to
f = clone_to(make_field(seed=1), "cuda")
for B in (1024, 8192):
    y0 = torch.randn(B, 16, device="cuda", requires_grad=True)
    t = torch.tensor([0.0, 1.0])
    for _ in range(2):
        sol = gode.odeint_adjoint(f, y0, t)
        sol.sum().backward()
    torch.cuda.synchronize()
    print("B", B, flush=True)
