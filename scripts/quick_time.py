"""Developer timing probe (not the bench): CUDA-event times of each fused kernel at a few batch sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gan_ode_b200 as gode
from tests.helpers import make_field, clone_to

dev = "cuda"
f = clone_to(make_field(seed=0), dev)
t = torch.linspace(0, 1, 16).float()


def timeit(fn, n=20, warm=5):
    """GPU time of fn's kernels: the device is kept busy by a spin kernel while the host enqueues, so the two
    events bracket device work only (no host launch gaps)."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in evs:
        torch.cuda._sleep(1000000)
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2] * 1e3  # us


for B in (16, 1024, 4096, 65536, 1048576):
    y0 = torch.randn(B, 16, device=dev)
    g = torch.randn(16, B, 16, device=dev)
    with torch.no_grad():
        t_f = timeit(lambda: gode.odeint(f, y0, t, method="rk4"))
    y0r = y0.clone().requires_grad_(True)
    sol = gode.odeint_adjoint(f, y0r, t, method="rk4")
    t_a = timeit(lambda: torch.autograd.grad(sol, [y0r] + list(f.parameters()), g, retain_graph=True))
    sol2 = gode.odeint(f, y0r, t, method="rk4")
    t_b = timeit(lambda: torch.autograd.grad(sol2, [y0r] + list(f.parameters()), g, retain_graph=True))
    with torch.no_grad():
        t_tf = timeit(lambda: gode.odeint(f, y0, t, method="rk4", options={"precision": "tf32"}))
        t_bf = timeit(lambda: gode.odeint(f, y0, t, method="rk4", options={"precision": "bf16"}))
    print("B=%8d rk4 fwd fp32 %9.1f us | tf32 %9.1f us | bf16 %9.1f us  -> fwd traj-steps/s fp32 %.3e tf32 %.3e bf16 %.3e" % (
        B, t_f, t_tf, t_bf, B * 15 / (t_f * 1e-6), B * 15 / (t_tf * 1e-6), B * 15 / (t_bf * 1e-6)), flush=True)
    line = "B=%8d rk4 fwd %9.1f us  adjoint bwd %9.1f us  backprop bwd %9.1f us | fwd+adj %.3e traj-steps/s" % (
        B, t_f, t_a, t_b, B * 15 / ((t_f + t_a) * 1e-6))
    if B <= 4096:
        with torch.no_grad():
            t_d = timeit(lambda: gode.odeint(f, y0, t, method="dopri5", rtol=1e-5, atol=1e-5))
        sol3 = gode.odeint(f, y0r, t, method="dopri5", rtol=1e-5, atol=1e-5)
        log = gode.last_step_log()
        t_db = timeit(lambda: torch.autograd.grad(sol3, [y0r] + list(f.parameters()), g, retain_graph=True))
        line += " | dopri5 fwd %8.1f us bwd %8.1f us (att %d acc %d)" % (t_d, t_db, log.n_attempts, log.n_accepted)
    print(line, flush=True)
