"""ODE-RNN sampler (configs[2]: B = 8192, 16 frames, torchdiffeq default tolerances): persistent all-frames forward kernel vs the
round-1 per-frame launches (GODE_ODERNN_PERFRAME=1); forward only and forward + backward, CUDA events, median of 9."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gan_ode_b200 as gode
from gan_ode_b200.fields import make_field

dev = "cuda"
f = make_field(16, 16, seed=0).to(dev)
cell = torch.nn.GRUCell(16, 16).to(dev)
B, F = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 16
h0 = torch.randn(B, 16, device=dev, requires_grad=True)
eps = torch.randn(F, B, 16, device=dev)
w = torch.randn(F, B, 16, device=dev)
params = list(f.parameters()) + list(cell.parameters())


def timeit(fn, n=9):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        torch.cuda._sleep(2000000)
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return round(sorted(a.elapsed_time(b) for a, b in ev)[n // 2] * 1e3, 1)


out = {"B": B, "F": F}
for perframe in ("1", "0"):
    os.environ["GODE_ODERNN_PERFRAME"] = perframe
    tag = "per_frame_launches" if perframe == "1" else "persistent_kernel"

    def fwd():
        with torch.no_grad():
            return gode.odernn_codes(f, cell, h0, eps)

    r = {"fwd_us": timeit(fwd)}
    for mode in ("discrete", "continuous"):
        def both():
            codes = gode.odernn_codes(f, cell, h0, eps, options={"adjoint": mode})
            torch.autograd.grad((codes * w).sum(), [h0] + params)
        r["fwd_bwd_%s_us" % mode] = timeit(both)
    out[tag] = r
print(json.dumps(out))
