"""Developer timing probe for the wide field D=64, H=256: FP32 warp-per-trajectory vs tcgen05 BF16 forward, FP32 adjoint."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gan_ode_b200 as gode
from tests.helpers import make_field, clone_to

dev = "cuda"
f = clone_to(make_field(64, 256, seed=0), dev)
t = torch.linspace(0, 1, 16).float()


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in evs:
        torch.cuda._sleep(1000000)
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2] * 1e3


only_tc = len(sys.argv) > 1 and sys.argv[1] == "tc"
for B in ((16384,) if only_tc else (128, 2048, 4096, 18944, 65536)):
    y0 = torch.randn(B, 64, device=dev)
    with torch.no_grad():
        t_bf = timeit(lambda: gode.odeint(f, y0, t, method="rk4", options={"precision": "bf16"}))
        if only_tc:
            continue
        t_f = timeit(lambda: gode.odeint(f, y0, t, method="rk4"))
    flops = B * 15 * 262144
    line = "B=%6d wide rk4 fwd fp32 %9.1f us (%.1f TFLOP/s) | tcgen05 bf16 %9.1f us (%.1f TFLOP/s, %.3e traj-steps/s)" % (
        B, t_f, flops / t_f * 1e-6, t_bf, flops / t_bf * 1e-6, B * 15 / (t_bf * 1e-6))
    if B <= 4096:
        g = torch.randn(16, B, 64, device=dev)
        y0r = y0.clone().requires_grad_(True)
        sol = gode.odeint_adjoint(f, y0r, t, method="rk4")
        t_a = timeit(lambda: torch.autograd.grad(sol, [y0r] + list(f.parameters()), g, retain_graph=True), n=5, warm=2)
        line += " | fp32 adjoint bwd %9.1f us" % t_a
    print(line, flush=True)
print("done")
