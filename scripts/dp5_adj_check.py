"""Developer check: continuous dopri5 adjoint kernel vs the oracle's odeint_adjoint (CPU)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gan_ode_b200 as gode
from oracle import torchdiffeq_restatement as tdq
from tests.helpers import make_field, clone_to, rel_err

DEV = "cuda"
for (B, t, scale, tol) in [(24, torch.tensor([0.0, 1.0]), 1.0, (1e-7, 1e-9)), (1, torch.tensor([0.0, 1.0]), 2.0, (1e-5, 1e-5)),
                           (37, torch.tensor([0.0, 0.4, 1.0]), 2.0, (1e-6, 1e-8)), (300, torch.tensor([1.0, 0.5, 0.0]), 3.0, (1e-5, 1e-6)),
                           (2048, torch.linspace(0, 1, 5), 2.0, (1e-5, 1e-5))]:
    f = make_field(seed=B, scale=scale)
    torch.manual_seed(B)
    y0 = torch.randn(B, 16)
    g = torch.randn(len(t), B, 16)

    def run(fn, field, y, gg, **kw):
        y = y.clone().requires_grad_(True)
        sol = fn(field, y, t, rtol=tol[0], atol=tol[1], **kw)
        return torch.autograd.grad((sol * gg).sum(), [y] + list(field.parameters()))

    t0 = time.time()
    ref = run(tdq.odeint_adjoint, f, y0, g)
    rl = tdq.last_step_log()
    t1 = time.time()
    out = run(gode.odeint_adjoint, clone_to(f, DEV), y0.to(DEV), g.to(DEV))
    al = gode.last_adjoint_log()
    dis = run(gode.odeint_adjoint, clone_to(f, DEV), y0.to(DEV), g.to(DEV), options={"adjoint": "discrete"})
    print("B", B, "T", len(t), "oracle %.1fs" % (t1 - t0), "status", al.status, "att", al.n_attempts, "acc", al.n_accepted, "nfe", al.nfe)
    print("  cont vs oracle:", ["%.2e" % rel_err(a, b) for a, b in zip(out, ref)])
    print("  disc vs oracle:", ["%.2e" % rel_err(a, b) for a, b in zip(dis, ref)])
    n_last = len(rl.accepted)
    print("  oracle last interval: acc", rl.accepted, "dt0 %.4g" % rl.dt0)
    print("  kernel tail:         acc", al.accepted[-n_last:], "dt0 %.4g" % al.dt0)
    print("  er oracle", ["%.3g" % e for e in rl.error_ratio][:8]); print("  er kernel", ["%.3g" % e for e in al.error_ratio[-n_last:]][:8])

# timing at the ODE-RNN shape
f = clone_to(make_field(seed=1), DEV)
for B in (1024, 8192):
    y0 = torch.randn(B, 16, device=DEV, requires_grad=True)
    t = torch.tensor([0.0, 1.0])
    for mode in ("continuous", "discrete"):
        for _ in range(3):
            sol = gode.odeint_adjoint(f, y0, t, options={"adjoint": mode})
            sol.sum().backward()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sol = gode.odeint_adjoint(f, y0, t, options={"adjoint": mode})
        s.record(); sol.sum().backward(); e.record(); torch.cuda.synchronize()
        extra = ""
        if mode == "continuous":
            al = gode.last_adjoint_log()
            extra = "attempts %d" % al.n_attempts
        print("B", B, mode, "backward %.3f ms" % s.elapsed_time(e), extra)
