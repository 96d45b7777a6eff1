"""Developer probe: one tcgen05 BF16 wide-field (D=64, H=256) rk4 forward at batch B (argv[1], default 37888 = 2 tiles/SM),
timed with CUDA events; run under ncu for the profile."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gan_ode_b200 as gode
from tests.helpers import make_field, clone_to

B = int(sys.argv[1]) if len(sys.argv) > 1 else 37888
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10
f = clone_to(make_field(64, 256, seed=0), "cuda")
t = torch.linspace(0, 1, 16).float()
y0 = torch.randn(B, 64, device="cuda")
ts = []
with torch.no_grad():
    for i in range(n + 3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); gode.odeint(f, y0, t, method="rk4", options={"precision": "bf16"}); b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
ms = sorted(ts)[len(ts) // 2]
fl = B * 15 * 262144
print("B=%d  %.1f us  %.1f TFLOP/s  %.3e traj-steps/s  (%.1f%% of 1348.6 TF sustained bf16 GEMM)" % (
    B, ms * 1e3, fl / ms * 1e-9, B * 15 / ms * 1e3, fl / ms * 1e-9 / 1348.6 * 100))
