import os, sys, torch
sys.path.insert(0, os.getcwd())
import gan_ode_b200 as gode
from tests.helpers import SDEFunc
torch.manual_seed(0)
sde = SDEFunc(16, 16).cuda()
ts = torch.linspace(0, 1, 16).float()
for B in (16384, 131072, 524288):
    y0 = torch.randn(B, 16, device="cuda", requires_grad=True)
    gr = torch.randn(16, B, 16, device="cuda")
    params = list(sde.parameters())
    def fwd():
        return gode.sdeint(sde, y0, ts, method="euler", dt=2.5e-2, bm=gode.PhiloxBrownian(1234, 0))
    sol = fwd()
    def bwd():
        return torch.autograd.grad(sol, [y0] + params, gr, retain_graph=True)
    def t(fn):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / 5
    with torch.no_grad():
        tf = t(fwd)
    tb = t(bwd)
    print("B=%d SDE fwd %.3f ms bwd %.3f ms -> %.3e traj-steps/s fwd+bwd (%.1f TFLOP/s of 24DH)" % (B, tf, tb, B * 41 / (tf + tb) * 1e3, B * 41 * 6144 / (tf + tb) * 1e-9))
