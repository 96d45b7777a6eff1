"""Headline-step probe (configs[1]: B=4096, dopri5 1e-5, fwd + backprop): graph-replay time of the forward alone and of the
step with / without the programmatic dependent launch of the backward; eager step; reference-shape rk4 + adjoint likewise.
    python scripts/headline_probe.py [B]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gan_ode_b200 as gode
from gan_ode_b200.fields import make_field

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
f = make_field(16, 16, seed=0).to(dev)
params = list(f.parameters())
g = torch.Generator().manual_seed(1000)
y0 = torch.randn(B, 16, generator=g).to(dev).requires_grad_(True)
grad = torch.randn(16, B, 16, generator=g).to(dev)
t = torch.linspace(0, 1, 16).float()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def graph_of(fn, pdl):
    gode.config.pdl = pdl
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        keep = fn()
    gode.config.pdl = False
    return gr, keep


def replay_us(gr, k=60):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
    for _ in range(5):
        gr.replay()
    for a, b in evs:
        flush.fill_(1)
        a.record(); gr.replay(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return round(ts[k // 2] * 1e3, 2), round(ts[0] * 1e3, 2)


out = {"B": B}
for name, kw, solve in (("dopri5_backprop", dict(method="dopri5", rtol=1e-5, atol=1e-5), gode.odeint),
                        ("rk4_adjoint", dict(method="rk4"), gode.odeint_adjoint)):
    def fwd():
        with torch.no_grad():
            return solve(f, y0, t, **kw)

    def step():
        sol = solve(f, y0, t, **kw)
        return torch.autograd.grad(sol, [y0] + params, grad)

    ref = [x.clone() for x in step()]
    r = {}
    gf, _ = graph_of(fwd, False)
    r["fwd_graph_us(median,min)"] = replay_us(gf)
    for warps in ("4", "7"):
        os.environ["GODE_DP5_BWD_WARPS"] = warps
        for pdl in (False, True):
            gs, keep = graph_of(step, pdl)
            r["step_graph_bwdwarps=%s_pdl=%d_us" % (warps, pdl)] = replay_us(gs)
            gs.replay()
            torch.cuda.synchronize()
            r["bwdwarps=%s_pdl=%d_maxrel_vs_eager" % (warps, pdl)] = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(keep, ref))
    os.environ.pop("GODE_DP5_BWD_WARPS")
    # eager issue rate (host-bound)
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(200):
        step()
    torch.cuda.synchronize()
    r["eager_step_us"] = round((time.perf_counter() - t0) / 200 * 1e6, 1)
    out[name] = r
print(json.dumps(out))
