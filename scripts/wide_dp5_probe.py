"""Wide-field dopri5: step logs of the CUDA solve beside the CPU oracle's (developer tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gan_ode_b200 as gode
from oracle import torchdiffeq_restatement as tdq
from tests.helpers import clone_to, make_field

t = torch.linspace(0, 1, 16)
for (D, H, B, scale, seed) in [(32, 32, 1, 1.0, 0), (32, 32, 8, 1.0, 0), (32, 32, 64, 1.0, 0), (32, 64, 37, 2.0, 0),
                               (64, 256, 19, 1.0, 0), (64, 256, 300, 2.0, 0), (16, 16, 1, 1.0, 0), (16, 16, 8, 1.0, 0)]:
    f = make_field(D, H, seed=seed, scale=scale)
    torch.manual_seed(seed + 1)
    y0 = torch.randn(B, D)
    with torch.no_grad():
        out = gode.odeint(clone_to(f, "cuda"), y0.cuda(), t, method="dopri5", rtol=1e-5, atol=1e-5)
        g = gode.last_step_log()
        ref = tdq.odeint(f, y0, t, method="dopri5", rtol=1e-5, atol=1e-5, options={"_replay_dt": g.dt})
        r = tdq.last_step_log()
        r64o = tdq.odeint(clone_to(f, "cpu", torch.float64), y0.double(), t.double(), method="dopri5", rtol=1e-5, atol=1e-5,
                          options={"_replay_dt": g.dt})
        r64 = tdq.last_step_log()
    print(D, H, B, scale, "dt", ["%.5g" % x for x in g.dt])
    print("   er gpu   ", ["%.6g" % x for x in g.error_ratio])
    print("   er cpu32 ", ["%.6g" % x for x in r.error_ratio])
    print("   er cpu64 ", ["%.6g" % x for x in r64.error_ratio])
    print("   traj rel err vs cpu32 %.3g  vs cpu64 %.3g ; cpu32 vs cpu64 %.3g" % (
        float((out.cpu() - ref).norm() / ref.norm()), float((out.cpu().double() - r64o).norm() / r64o.norm()),
        float((ref.double() - r64o).norm() / r64o.norm())))
