"""Developer timing probe: ODE-RNN sampler (configs[2]: B=8192, 16 frames, torchdiffeq default tolerances), fused C-level
loop + GRU kernel vs the reference loop through the shim (16 dopri5 launches + nn.GRUCell in PyTorch), fwd and fwd+bwd."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gan_ode_b200 as gode
from tests.caller_model import LatentMotionODERNN

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
F = 16
torch.manual_seed(0)
m = LatentMotionODERNN(16, F).cuda()
h0 = torch.randn(B, 16, device="cuda")
eps = torch.randn(F, B, 16, device="cuda")
w = torch.randn(B * F, 16, device="cuda")


def timeit(fn, n=7):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(2000000)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[n // 2]


def fused(bwd, mode="discrete"):
    m.zero_grad()
    codes = gode.odernn_codes(m.ode_fn, m.recurrent, h0, eps, options={"adjoint": mode})
    if bwd:
        (codes.transpose(0, 1).reshape(-1, 16) * w).sum().backward()


def unfused(bwd):
    m.zero_grad()
    out = m.sample_z_m(B, h0=h0, eps=eps)
    if bwd:
        (out * w).sum().backward()


gode.install_shims()
with torch.no_grad():
    tf0, tu0 = timeit(lambda: fused(False)), timeit(lambda: unfused(False))
tf1, tu1 = timeit(lambda: fused(True)), timeit(lambda: unfused(True))
tfc = timeit(lambda: fused(True, "continuous"))
gode.config.dopri5_adjoint = "discrete"
tud = timeit(lambda: unfused(True))
gode.config.dopri5_adjoint = "continuous"
fr = gode.odernn.last_log().frames()
att = sum(f["n_attempts"] for f in fr)
print("B=%d F=%d attempted steps (sum over frames) %d, accepted %d" % (B, F, att, sum(f["n_accepted"] for f in fr)))
print("forward      fused %.3f ms | shim loop %.3f ms" % (tf0, tu0))
print("fwd+bwd      fused %.3f ms | shim loop %.3f ms  -> %.3e trajectory-steps/s fused   (discrete gradient | continuous adjoint)" % (tf1, tu1, B * att / tf1 * 1e3))
print("fwd+bwd      fused, continuous adjoint %.3f ms | shim loop, discrete gradient %.3f ms" % (tfc, tud))
