"""ncu target: a few launches of the tensor-core and FP32 rk4 forward at a large batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gan_ode_b200 as gode
from tests.helpers import make_field, clone_to
B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
f = clone_to(make_field(seed=0), "cuda")
t = torch.linspace(0, 1, 16).float()
y0 = torch.randn(B, 16, device="cuda")
with torch.no_grad():
    for prec in ("fp32", "tf32", "bf16", "fp32", "tf32", "bf16"):
        gode.odeint(f, y0, t, method="rk4", options={"precision": prec})
torch.cuda.synchronize()
print("ok")
