"""Developer timing: continuous dopri5 adjoint backward, 8 vs 4 lanes per trajectory (GODE_ADJ_LANES), two-point grid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gan_ode_b200 as gode
from tests.helpers import make_field, clone_to

f = clone_to(make_field(seed=1), "cuda")
t = torch.tensor([0.0, 1.0])
for B in (256, 1024, 2048, 4096, 8192):
    y0 = torch.randn(B, 16, device="cuda", requires_grad=True)
    row = []
    ref = None
    for lanes in ("8", "4"):
        os.environ["GODE_ADJ_LANES"] = lanes
        ts = []
        for it in range(8):
            sol = gode.odeint_adjoint(f, y0, t)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); g = torch.autograd.grad(sol.sum(), [y0] + list(f.parameters())); e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        n = gode.last_adjoint_log().n_attempts
        if ref is None:
            ref = g
        else:
            row.append("max rel diff %.1e" % max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(g, ref)))
        row.append("L=%s %.3f ms (%d attempts)" % (lanes, sorted(ts)[len(ts) // 2], n))
    print("B", B, " | ".join(row))
