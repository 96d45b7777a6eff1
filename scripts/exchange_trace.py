"""In-kernel timing of the fused rank exchange at the tail of dopri5_backprop_bwd_kernel (developer tool, trace build):
    python -m gan_ode_b200.build --trace
    GODE_LIB=gan_ode_b200/csrc/libgode_trace.so torchrun --nproc-per-node N ... scripts/exchange_trace.py
Every rank replays the captured step (forward + backward with the exchange) and prints, for the last replay, CTA 0 / thread 0's
stamps: row stored -> own columns reduced and published (local part) -> kernel exit (= all peers' words seen and added)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import gan_ode_b200 as gode
from gan_ode_b200 import _lib
from gan_ode_b200.fields import make_field

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    from gan_ode_b200.dist import enable_fused_grad_exchange
    assert enable_fused_grad_exchange(require=True)
B = 4096
f = make_field(16, 16, seed=0).to(dev)
params = list(f.parameters())
g = torch.Generator().manual_seed(1000 + rank)
y0 = torch.randn(B, 16, generator=g).to(dev).requires_grad_(True)
grad = torch.randn(16, B, 16, generator=g).to(dev)
t = torch.linspace(0, 1, 16).float()
kw = dict(method="dopri5", rtol=1e-5, atol=1e-5)
L = _lib.lib()
L.gode_debug_trace_read.argtypes = [C.c_void_p]
L.gode_debug_trace_read.restype = C.c_int


def step():
    return torch.autograd.grad(gode.odeint(f, y0, t, **kw), [y0] + params, grad)


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
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(200):
    gr.replay()
b.record()
torch.cuda.synchronize()
buf = (C.c_ulonglong * 256)()
assert L.gode_debug_trace_read(buf) == 0
bw = [(buf[(64 + s_) * 2], buf[(64 + s_) * 2 + 1]) for s_ in range(64)]
t0 = bw[0][0]
line = "rank %d: %.2f us/step | bwd entry 0, replay done %.2f, row stored %.2f, own columns published %.2f, exit %.2f  => exchange wait %.2f us" % (
    rank, a.elapsed_time(b) * 5, (bw[30][0] - t0) / 1e3, (bw[33][0] - t0) / 1e3, (bw[34][0] - t0) / 1e3, (bw[31][0] - t0) / 1e3,
    (bw[31][0] - bw[34][0]) / 1e3)
for r in range(world):
    if r == rank:
        print(line, flush=True)
    if world > 1:
        dist.barrier()
if world > 1:
    torch.cuda.synchronize()
    os._exit(0)
