"""Launch-gap probe: the headline step (dopri5 forward + backprop, B = 4096) captured as graphs of 1 / 2 / 4 / 8 steps over
distinct resident batches, replayed back to back; µs per STEP.  Inside a graph the kernel -> kernel gap is ~2 µs, between two
graph launches on one stream it is larger."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gan_ode_b200 as gode
from gan_ode_b200.fields import make_field

dev = torch.device("cuda", 0)
B, NB = 4096, 24
f = make_field(16, 16, seed=0).to(dev)
params = list(f.parameters())
t = torch.linspace(0, 1, 16).float()
kw = dict(method="dopri5", rtol=1e-5, atol=1e-5)
batches = []
for i in range(NB):
    g = torch.Generator().manual_seed(1000 + 7919 * (i + 1))
    batches.append((torch.randn(B, 16, generator=g).to(dev).requires_grad_(True), torch.randn(16, B, 16, generator=g).to(dev)))


def step(y, gr):
    return torch.autograd.grad(gode.odeint(f, y, t, **kw), [y] + params, gr)


for y, gr in batches[:2]:
    step(y, gr)
torch.cuda.synchronize()
out = {}
for per in (1, 2, 4, 8):
    graphs, keep = [], []
    for g0 in range(0, NB, per):
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            step(*batches[g0])
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        gph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph):
            keep.append([step(*batches[g0 + k]) for k in range(per)])
        graphs.append(gph)
    for gph in graphs:
        gph.replay()
    torch.cuda.synchronize()
    reps = 480 // NB
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        for gph in graphs:
            gph.replay()
    b.record()
    torch.cuda.synchronize()
    out["steps_per_graph=%d" % per] = round(a.elapsed_time(b) * 1e3 / (reps * NB), 2)
    del graphs, keep
print(json.dumps(out))
