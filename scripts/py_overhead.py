"""Developer probe: host-side cost of the eager public API at the reference's real batch sizes (B=32 / 1024, rk4 + adjoint):
wall time per fwd+bwd call and a cProfile of where the Python time goes."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gan_ode_b200 as gode
from tests.helpers import make_field, clone_to

f = clone_to(make_field(16, 16, seed=0), "cuda")
t = torch.linspace(0, 1, 16).float()
params = list(f.parameters())
for B in (32, 1024):
    y0 = torch.randn(B, 16, device="cuda", requires_grad=True)
    g = torch.randn(16, B, 16, device="cuda")

    def step():
        sol = gode.odeint_adjoint(f, y0, t, method="rk4")
        return torch.autograd.grad(sol, [y0] + params, g)

    for _ in range(20):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(300):
        step()
    torch.cuda.synchronize()
    print("B=%d rk4+adjoint eager fwd+bwd: %.1f us per call (wall, 300 calls back to back)" % (B, (time.perf_counter() - t0) / 300 * 1e6))
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
