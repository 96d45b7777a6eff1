"""Where the host time of one EAGER step goes (developer tool): the bench's e2e_eager_step under cProfile, plus wall-clock
splits of its phases with the GPU idle-waited between them.   python scripts/eager_profile.py [B]"""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gan_ode_b200 as gode
from gan_ode_b200.fields import make_field

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
f = make_field(16, 16, seed=0).to(dev)
params = list(f.parameters())
g = torch.Generator().manual_seed(1000)
y0_host = torch.randn(B, 16, generator=g).pin_memory()
grad = torch.randn(16, B, 16, generator=g).to(dev)
t = torch.linspace(0, 1, 16).float()
kw = dict(method="dopri5", rtol=1e-5, atol=1e-5)
res_host = torch.empty(sum(p.numel() for p in params)).pin_memory()


def step():
    y = y0_host.to(dev, non_blocking=True).requires_grad_(True)
    sol = gode.odeint(f, y, t, **kw)
    gs_ = torch.autograd.grad(sol, params, grad)
    res_host.copy_(torch.cat([x.reshape(-1) for x in gs_]), non_blocking=True)
    torch.cuda.current_stream().synchronize()


for _ in range(20):
    step()
N = 300
t0 = time.perf_counter()
for _ in range(N):
    step()
print("eager step: %.1f us" % ((time.perf_counter() - t0) / N * 1e6))

# phase split: host time to ISSUE each phase (no sync in between), then the final wait
acc = [0.0] * 5
for _ in range(N):
    a = time.perf_counter()
    y = y0_host.to(dev, non_blocking=True).requires_grad_(True)
    b = time.perf_counter()
    sol = gode.odeint(f, y, t, **kw)
    c = time.perf_counter()
    gs_ = torch.autograd.grad(sol, params, grad)
    d = time.perf_counter()
    res_host.copy_(torch.cat([x.reshape(-1) for x in gs_]), non_blocking=True)
    e = time.perf_counter()
    torch.cuda.current_stream().synchronize()
    f_ = time.perf_counter()
    for i, v in enumerate((b - a, c - b, d - c, e - d, f_ - e)):
        acc[i] += v
print("issue H2D %.1f | odeint %.1f | autograd.grad %.1f | cat + D2H %.1f | final sync wait %.1f  (us)" %
      tuple(x / N * 1e6 for x in acc))

# the same with autograd's worker threads off: the CUDA backward node then runs on the calling thread (no hand-off to the
# engine's device thread and back)
with torch.autograd.set_multithreading_enabled(False):
    for _ in range(20):
        step()
    t0 = time.perf_counter()
    for _ in range(N):
        step()
    print("eager step, torch.autograd.set_multithreading_enabled(False): %.1f us" % ((time.perf_counter() - t0) / N * 1e6))

pr = cProfile.Profile()
pr.enable()
for _ in range(N):
    step()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print(s.getvalue()[:6000])
