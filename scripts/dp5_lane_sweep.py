import torch, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gan_ode_b200 as gode
from tests.helpers import make_field, clone_to
f = clone_to(make_field(seed=0), "cuda"); t = torch.linspace(0,1,16).float()
for B in (32, 256, 1024, 2048, 4096, 7000):
    torch.manual_seed(B)
    y = torch.randn(B,16,device="cuda")
    g = torch.cuda.CUDAGraph()
    with torch.no_grad():
        for _ in range(3): gode.odeint(f,y,t,method="dopri5",rtol=1e-5,atol=1e-5)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            gode.odeint(f,y,t,method="dopri5",rtol=1e-5,atol=1e-5)
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): g.replay()
    b.record(); torch.cuda.synchronize()
    print("L=%s B=%d dopri5 fwd %.1f us (graph replay, back to back)" % (os.environ.get("GODE_DP5_LANES","8"), B, a.elapsed_time(b)*1e3/50))
