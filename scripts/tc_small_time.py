"""Developer timing probe: bf16 tcgen05 rk4 forward at D=H=16, B = 1M."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gan_ode_b200 as gode
from tests.helpers import make_field, clone_to
f = clone_to(make_field(16, 16, seed=0), "cuda")
t = torch.linspace(0, 1, 16).float()
for B in (4096, 1 << 18, 1 << 20):
    y0 = torch.randn(B, 16, device="cuda")
    with torch.no_grad():
        for _ in range(3): gode.odeint(f, y0, t, method="rk4", options={"precision": "bf16"})
        torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            torch.cuda._sleep(1000000)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); gode.odeint(f, y0, t, method="rk4", options={"precision": "bf16"}); b.record()
            torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ms = sorted(ts)[3]
    print("B=%d bf16 rk4 fwd %.3f ms  %.3e traj-steps/s  (%.1f%% of HBM roofline at 64 B/traj-step)" % (B, ms, B * 15 / ms * 1e3, B * 15 * 64 / ms * 1e3 / 6533.8e9 * 100))
