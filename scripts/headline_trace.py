"""In-kernel timeline of the headline step (developer tool; needs the trace build):
    python -m gan_ode_b200.build --trace && GODE_LIB=gan_ode_b200/csrc/libgode_trace.so python scripts/headline_trace.py
Thread 0 of CTA 0 stamps (%globaltimer, clock64) inside dopri5_fwd_kernel / dopri5_backprop_bwd_kernel; this replays the captured
step (with and without the programmatic dependent launch of the backward) and prints the stamps of the LAST replay in µs
relative to the forward's entry (global timer) and the per-phase cycle counts (clock64, CTA 0's SM)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gan_ode_b200 as gode
from gan_ode_b200 import _lib
from gan_ode_b200.fields import make_field

B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4096
dev = torch.device("cuda", 0)
f = make_field(16, 16, seed=0).to(dev)
params = list(f.parameters())
g = torch.Generator().manual_seed(1000)
y0 = torch.randn(B, 16, generator=g).to(dev).requires_grad_(True)
grad = torch.randn(16, B, 16, generator=g).to(dev)
t = torch.linspace(0, 1, 16).float()
kw = dict(method="dopri5", rtol=1e-5, atol=1e-5)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
L = _lib.lib()
L.gode_debug_trace_read.argtypes = [C.c_void_p]
L.gode_debug_trace_read.restype = C.c_int


def step():
    sol = gode.odeint(f, y0, t, **kw)
    return torch.autograd.grad(sol, [y0] + params, grad)


def read():
    buf = (C.c_ulonglong * 256)()
    assert L.gode_debug_trace_read(buf) == 0
    return [[(buf[(k * 64 + s) * 2], buf[(k * 64 + s) * 2 + 1]) for s in range(64)] for k in range(2)]


out = {}
for pdl in (False, True):
    gode.config.pdl = pdl
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        step()
    gode.config.pdl = False
    for _ in range(5):
        if "--warm" not in sys.argv:
            flush.fill_(1)
        gr.replay()
    torch.cuda.synchronize()
    fw, bw = read()
    t0 = fw[0][0]
    names_f = {0: "entry", 1: "weights+y0 loaded", 2: "f0", 3: "initial step chosen", 40: "exit"}
    for a in range(4):
        names_f[4 + 3 * a] = "att%d stages" % a
        names_f[5 + 3 * a] = "att%d reduced" % a
        names_f[6 + 3 * a] = "att%d end" % a
    names_b = {0: "entry", 1: "row weights loaded", 2: "griddep_wait + log read", 50: "grad copies issued", 51: "column weights staged",
               3: "CTA barrier", 52: "accumulators cleared", 53: "first checkpoint load issued", 4: "grads staged", 30: "replay done", 32: "reduce: warp sums in smem", 33: "reduce: CTA partial stored", 36: "reduce: LAST CTA's partial stored",
               34: "reduce: grid barrier passed", 31: "exit"}
    for st in range(4):
        names_b[5 + 2 * st] = "step%d begin" % st
        names_b[6 + 2 * st] = "step%d recomputed" % st
    rows = []
    for nm, tr, names in (("fwd", fw, names_f), ("bwd", bw, names_b)):
        prev_c = None
        for slot in sorted(names):
            gt, ck = tr[slot]
            if gt == 0:
                continue
            rows.append((nm, names[slot], round((gt - t0) / 1e3, 2), None if prev_c is None else int(ck - prev_c)))
            prev_c = ck
    rows.sort(key=lambda r: r[2])
    out["pdl=%d" % pdl] = rows
for k, rows in out.items():
    print(k)
    for r in rows:
        print("   %-4s %-22s t=%8.2f us   +%s cycles" % r)
