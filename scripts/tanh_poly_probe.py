"""Wide tensor-core forward with a share of the tanh evaluations on the FMA pipe (GODE_TANH_POLY_N = pairs of every eight, read once per process):
time at two batch sizes and error against the FP32 wide forward.   python scripts/tanh_poly_probe.py"""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) == 1:
    for pe in ("0", "1", "2", "3", "4"):
        env = dict(os.environ, GODE_TANH_POLY_N=pe)
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print(pe, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:])
    sys.exit(0)
import torch

import gan_ode_b200 as gode
from gan_ode_b200.fields import make_field

f = make_field(64, 256, seed=0).to("cuda")
t = torch.linspace(0, 1, 16)
out = {}
for B in (18944, 151552):
    y0 = torch.randn(B, 64, device="cuda")
    with torch.no_grad():
        fn = lambda: gode.odeint(f, y0, t, method="rk4", options={"precision": "bf16"})
        for _ in range(3):
            fn()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        torch.cuda.synchronize()
        for a, b in ev:
            a.record(); fn(); b.record()
        torch.cuda.synchronize()
        us = sorted(a.elapsed_time(b) for a, b in ev)[5] * 1e3
        out["B%d_us" % B] = round(us, 1)
        out["B%d_tflops" % B] = round(B * 15 * 262144 / us * 1e-6, 1)
        if B == 18944:
            ref = gode.odeint(f, y0[:2048], t, method="rk4")
            got = gode.odeint(f, y0[:2048], t, method="rk4", options={"precision": "bf16"})
            out["rel_err_vs_fp32"] = float((got - ref).norm() / ref.norm())
print(json.dumps(out))
