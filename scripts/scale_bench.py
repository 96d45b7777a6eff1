"""Multi-GPU side measurements of the BASELINE.json configs that bench.py's headline does not cover (one JSON line per
workload, rank 0).  Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1
--master-port 29511 scripts/scale_bench.py     (or plain `python scripts/scale_bench.py` for N = 1)

  odernn   configs[2]: ODE-RNN sampler, 16 frames, B = 8192 TOTAL sharded over the ranks (strong scaling), torchdiffeq
           default tolerances, fused C-level loop, fwd + bwd, gradients of both parameter sets all-reduced (NCCL)
  sde      configs[4]: Euler-Maruyama, in-kernel Philox keyed by GLOBAL trajectory index, B = 16384 TOTAL sharded, fwd + bwd
  wide     configs[3]'s shape: D=64/H=256 rk4 + tensor-core adjoint, B = 37888 PER GPU (weak), ODE gradients all-reduced
Timing: CUDA events around each step, barrier + synchronize on both sides, MAX over ranks; trajectory-steps counted
over ALL ranks.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import gan_ode_b200 as gode
from gan_ode_b200.dist import shard_bounds
from tests.caller_model import LatentMotionODERNN
from tests.helpers import SDEFunc, make_field

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
OUT_FD = 1
if world > 1:
    sys.stdout.flush()
    OUT_FD = os.dup(1)   # NCCL's banner goes to fd 1 from native code: keep the JSON lines on the real stdout only
    os.dup2(2, 1)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
    gode.config.grad_allreduce = True
STEPS, WARM = 10, 3


def timed(fn):
    for _ in range(WARM):
        fn()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(STEPS)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    tt = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return float(tt.item()) / STEPS  # ms per step


def emit(name, scaling, total_traj_steps, ms, **extra):
    if rank == 0:
        os.write(OUT_FD, (json.dumps(dict(workload=name, n_gpus=world, scaling=scaling, ms_per_step=ms,
                                          trajectory_steps_per_s=total_traj_steps / ms * 1e3, **extra)) + "\n").encode())


which = [a for a in sys.argv[1:] if a != "weak"] or ["odernn", "sde", "wide"]
WEAK = "weak" in sys.argv[1:]   # odernn / sde: fixed batch PER GPU instead of a fixed total
if "odernn" in which:
    torch.manual_seed(0)
    m = LatentMotionODERNN(16, 16).to(dev)
    NB = 8192 * world if WEAK else 8192
    lo, hi = shard_bounds(NB, rank, world)
    g = torch.Generator().manual_seed(1)
    h0 = torch.randn(NB, 16, generator=g)[lo:hi].to(dev)
    eps = torch.randn(16, NB, 16, generator=g)[:, lo:hi].contiguous().to(dev)
    w = torch.randn(16, NB, 16, generator=g)[:, lo:hi].contiguous().to(dev)
    params = list(m.ode_fn.parameters()) + list(m.recurrent.parameters())

    def step():
        codes = gode.odernn_codes(m.ode_fn, m.recurrent, h0, eps)
        return torch.autograd.grad(codes, params, w)

    ms = timed(step)
    att = torch.tensor([sum(f["n_attempts"] for f in gode.odernn.last_log().frames())], device=dev)
    if world > 1:
        dist.all_reduce(att, op=dist.ReduceOp.MAX)  # per-rank batch-global controllers may differ by a step
    emit("odernn_B8192_%s_16frames_dopri5_default_tol_fwd_bwd" % ("per_gpu" if WEAK else "total"), "weak" if WEAK else "strong",
         NB * int(att.item()), ms,
         attempted_steps_per_solve_sum=int(att.item()))

if "sde" in which:
    torch.manual_seed(0)
    sde = SDEFunc(16, 16).to(dev)
    NS = 16384 * world if WEAK else 16384
    lo, hi = shard_bounds(NS, rank, world)
    g = torch.Generator().manual_seed(2)
    y0 = torch.randn(NS, 16, generator=g)[lo:hi].to(dev).requires_grad_(True)
    gr = torch.randn(16, NS, 16, generator=g)[:, lo:hi].contiguous().to(dev)
    ts = torch.linspace(0, 1, 16).float()
    params = list(sde.parameters())

    def step():
        sol = gode.sdeint(sde, y0, ts, method="euler", dt=2.5e-2, bm=gode.PhiloxBrownian(1234, lo))
        return torch.autograd.grad(sol, [y0] + params, gr)

    ms = timed(step)
    emit("sde_em_B16384_%s_41steps_philox_fwd_bwd" % ("per_gpu" if WEAK else "total"), "weak" if WEAK else "strong", NS * 41, ms)

if "wide" in which:
    f = make_field(64, 256, seed=0).to(dev)
    Bw = 37888
    g = torch.Generator().manual_seed(3 + rank)
    y0 = torch.randn(Bw, 64, generator=g).to(dev).requires_grad_(True)
    gr = torch.randn(16, Bw, 64, generator=g).to(dev)
    t = torch.linspace(0, 1, 16).float()
    params = list(f.parameters())

    def step():
        sol = gode.odeint_adjoint(f, y0, t, method="rk4", options={"precision": "bf16"})
        return torch.autograd.grad(sol, [y0] + params, gr)

    ms = timed(step)
    emit("rk4_adjoint_bf16_D64_H256_B37888_per_gpu_fwd_bwd_grad_allreduce", "weak", world * Bw * 15, ms,
         tflops=world * Bw * 15 * 1048576 / ms * 1e-9)

if world > 1:
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)  # NCCL teardown after graph-free runs is fine, but keep the exit path identical to bench.py
