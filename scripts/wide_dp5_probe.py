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

# timing: D=64 / H=256, rtol = atol = 1e-5, 16 outputs; forward and forward + recorded-step backward
import time
for B in (1184, 4096, 16384):
    f = clone_to(make_field(64, 256, seed=0), "cuda")
    y0 = torch.randn(B, 64, device="cuda", requires_grad=True)
    g = torch.randn(16, B, 64, device="cuda")
    params = list(f.parameters())

    def fwd():
        with torch.no_grad():
            return gode.odeint(f, y0, t, method="dopri5", rtol=1e-5, atol=1e-5)

    def both():
        return torch.autograd.grad(gode.odeint(f, y0, t, method="dopri5", rtol=1e-5, atol=1e-5), [y0] + params, g)

    res = {}
    for name, fn in (("fwd", fwd), ("fwd_bwd", both)):
        for _ in range(3):
            fn()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        torch.cuda.synchronize()
        for a, b in ev:
            a.record(); fn(); b.record()
        torch.cuda.synchronize()
        res[name] = sorted(a.elapsed_time(b) for a, b in ev)[5]
    lg = gode.last_step_log()
    flop = B * lg.nfe * 4 * 64 * 256   # 2 mat-vecs of D*H MACs per field evaluation
    print({"B": B, "attempts": lg.n_attempts, "nfe": lg.nfe, "fwd_ms": round(res["fwd"], 3), "fwd_bwd_ms": round(res["fwd_bwd"], 3),
           "fwd_fp32_tflops": round(flop / res["fwd"] / 1e9, 2),
           "trajectory_steps_per_s_fwd_bwd": round(B * lg.n_attempts / res["fwd_bwd"] * 1e3)})
if "--cpu" in sys.argv:
    B = 4096
    f = make_field(64, 256, seed=0)
    y0 = torch.randn(B, 64, requires_grad=True)
    g = torch.randn(16, B, 64)
    t0 = time.perf_counter()
    sol = tdq.odeint(f, y0, t, method="dopri5", rtol=1e-5, atol=1e-5)
    torch.autograd.grad(sol, [y0] + list(f.parameters()), g)
    print({"cpu_oracle_B4096_fwd_bwd_ms": round((time.perf_counter() - t0) * 1e3, 1), "threads": torch.get_num_threads()})
