"""Developer check (torchrun, N >= 2): peer-memory one-shot all-reduce vs NCCL — equality, repeated epochs, inside a CUDA
graph, and latency of both."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from gan_ode_b200.dist import P2PAllReduce

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ar = P2PAllReduce(cap=65536)
ok = True
for n in (544, 33088, 1632, 7):
    for rep in range(5):
        torch.manual_seed(100 * rep + rank)
        x = torch.randn(n, device="cuda")
        ref = x.clone()
        dist.all_reduce(ref)
        out = ar(x.clone())
        torch.cuda.synchronize()
        err = float((out - ref).abs().max() / ref.abs().max())
        ok &= err < 1e-6
# graph replay
x = torch.randn(544, device="cuda")
src = x.clone()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(2):
        y = src.clone(); ar(y)
torch.cuda.synchronize(); dist.barrier()
with torch.cuda.graph(g):
    y = src.clone(); ar(y)
for _ in range(5):
    g.replay()
torch.cuda.synchronize()
ref = src.clone(); dist.all_reduce(ref)
ok &= float((y - ref).abs().max() / ref.abs().max()) < 1e-6


def lat(fn, k=200):
    for _ in range(20):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / k * 1e3


z = torch.randn(544, device="cuda")
l1, l2 = lat(lambda: ar(z)), lat(lambda: dist.all_reduce(z))
z = torch.randn(33088, device="cuda")
l3, l4 = lat(lambda: ar(z)), lat(lambda: dist.all_reduce(z))
if rank == 0:
    print("p2p all-reduce ok=%s  | 544 floats: p2p %.1f us, nccl %.1f us | 33088 floats: p2p %.1f us, nccl %.1f us (back-to-back launches, N=%d)"
          % (ok, l1, l2, l3, l4, world), flush=True)
dist.barrier()
os._exit(0)
