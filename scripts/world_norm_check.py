"""Multi-GPU check of dopri5 with the world-scope error norm (gode_dopri5_fwd_world).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        scripts/world_norm_check.py

Every rank solves its (ragged) shard of one batch with options={'norm': 'world'}; rank 0 also solves the whole batch alone
with torchdiffeq's batch-global norm.  Expected: identical accept/reject flags, dt sequences equal to rounding, shard
solutions and all-reduced parameter gradients equal to the single-process ones.  Prints one JSON line on rank 0."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import gan_ode_b200 as gode
from gan_ode_b200 import dist as gdist
from tests.helpers import make_field, clone_to, rel_err

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
out_fd = os.dup(1)
os.dup2(2, 1)
dist.init_process_group("nccl", device_id=dev)
gdist.enable_world_norm()
gode.config.grad_allreduce = True

B = 1000
f = clone_to(make_field(seed=7, scale=4.0), dev)
torch.manual_seed(3)
y_full = (torch.randn(B, 16) * torch.linspace(0.3, 2.0, B).view(-1, 1)).to(dev)   # shards of different stiffness
g_full = torch.randn(16, B, 16).to(dev)
t = torch.linspace(0, 1, 16)
cuts = [0] + [int(B * (r + 1) / world) + (7 if r % 2 == 0 and r + 1 < world else 0) for r in range(world)]
cuts[-1] = B
lo, hi = cuts[rank], cuts[rank + 1]
res = {}
for rep in range(3):     # consecutive launches: the exchange epochs continue across them
    y = y_full[lo:hi].clone().requires_grad_(True)
    sol = gode.odeint(f, y, t, method="dopri5", rtol=1e-5, atol=1e-5, options={"norm": "world", "check": rep == 0})
    log = gode.last_step_log()
    grads = torch.autograd.grad((sol * g_full[:, lo:hi]).sum(), [y] + list(f.parameters()))
torch.cuda.synchronize()
mine = dict(status=log.status, accepted=log.accepted, dt=log.dt, er=log.error_ratio)
gathered = [None] * world
dist.all_gather_object(gathered, mine)
ok = True
if rank == 0:
    gode.config.grad_allreduce = None
    yf = y_full.clone().requires_grad_(True)
    ref = gode.odeint(f, yf, t, method="dopri5", rtol=1e-5, atol=1e-5)
    rl = gode.last_step_log()
    rg = torch.autograd.grad((ref * g_full).sum(), [yf] + list(f.parameters()))
    local_only = gode.odeint(f, y_full[lo:hi], t, method="dopri5", rtol=1e-5, atol=1e-5)
    ll = gode.last_step_log()
    res = dict(world=world, shards=[cuts[r + 1] - cuts[r] for r in range(world)], status=[m["status"] for m in gathered],
               n_attempts=len(rl.accepted), n_rejected=rl.n_rejected,
               flags_equal=all(m["accepted"] == rl.accepted for m in gathered),
               ranks_identical=all(m["dt"] == gathered[0]["dt"] and m["er"] == gathered[0]["er"] for m in gathered),
               dt_rel_diff=max(abs(a - b) / b for a, b in zip(gathered[0]["dt"], rl.dt)),
               er_rel_diff=max(abs(a - b) / max(b, 5e-2) for a, b in zip(gathered[0]["er"], rl.error_ratio)),
               sol_err=rel_err(sol, ref[:, lo:hi]), grad_y0_err=rel_err(grads[0], rg[0][lo:hi]),
               param_grad_err=max(rel_err(a, b) for a, b in zip(grads[1:], rg[1:])),
               per_rank_norm_would_differ=(ll.accepted != rl.accepted) or max(abs(a - b) / b for a, b in zip(ll.dt, rl.dt)) > 1e-3)
    ok = (res["flags_equal"] and res["ranks_identical"] and res["dt_rel_diff"] < 1e-3 and res["sol_err"] < 2e-5 and
          res["param_grad_err"] < 1e-4 and all(s == 0 for s in res["status"]))
    res["ok"] = ok
gode.config.grad_allreduce = None


# ---- parameter-gradient all-reduce fused into the backward kernel (gode_dopri5_backprop_bwd_world) vs ncclAllReduce -------
def grads_with(mode, solver="dopri5"):
    gode.config.grad_allreduce = True if mode == "nccl" else (p2p if mode == "p2p" else None)
    gode.config.grad_exchange = fused_ex if mode == "fused" else None
    outs = []
    for rep in range(3):
        yy = y_full[lo:hi].clone().requires_grad_(True)
        if solver == "dopri5":
            so = gode.odeint(f, yy, t, method="dopri5", rtol=1e-5, atol=1e-5)
        elif solver == "rk4_adjoint":
            so = gode.odeint_adjoint(f, yy, t, method="rk4")
        else:
            so = gode.odeint(f, yy, t, method="rk4")
        outs.append(torch.autograd.grad((so * g_full[:, lo:hi]).sum(), [yy] + list(f.parameters())))
    torch.cuda.synchronize()
    gode.config.grad_allreduce = gode.config.grad_exchange = None
    return outs


assert gdist.enable_fused_grad_exchange()
fused_ex = gode.config.grad_exchange
gode.config.grad_exchange = None
assert gdist.enable_p2p_allreduce()
p2p = gode.config.grad_allreduce
gode.config.grad_allreduce = None
rk4_err = {}
for solver in ("rk4_adjoint", "rk4_backprop"):
    a_, b_, c_ = grads_with("nccl", solver), grads_with("fused", solver), grads_with("p2p", solver)
    rk4_err[solver] = dict(fused_vs_nccl=max(rel_err(x, y_) for x, y_ in zip(b_[-1][1:], a_[-1][1:])),
                           p2p_vs_nccl=max(rel_err(x, y_) for x, y_ in zip(c_[-1][1:], a_[-1][1:])),
                           repeatable=all(torch.equal(x, y_) for x, y_ in zip(b_[0][1:], b_[-1][1:])))
g_nccl, g_fused = grads_with("nccl"), grads_with("fused")
flat = torch.cat([x.reshape(-1) for x in g_fused[-1][1:]])
gathered_g = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(gathered_g, flat)
if rank == 0:
    fe = dict(vs_nccl=max(rel_err(a, b) for a, b in zip(g_fused[-1][1:], g_nccl[-1][1:])),
              repeatable=all(torch.equal(a, b) for a, b in zip(g_fused[0][1:], g_fused[-1][1:])),
              identical_on_all_ranks=all(torch.equal(x, gathered_g[0]) for x in gathered_g),
              grad_y0_local=bool(torch.equal(g_fused[-1][0], g_nccl[-1][0])))
    res["fused_grad_exchange"] = fe
    res["rk4_grad_exchange"] = rk4_err
    res["ok"] = ok = bool(ok and fe["vs_nccl"] < 1e-5 and fe["repeatable"] and fe["identical_on_all_ranks"] and fe["grad_y0_local"]
                          and all(v["fused_vs_nccl"] < 1e-5 and v["p2p_vs_nccl"] < 1e-5 and v["repeatable"] for v in rk4_err.values()))


def timed(opts, Bt=4096, n=20):
    yt = torch.randn(Bt, 16, device=dev)
    with torch.no_grad():
        for _ in range(3):
            gode.odeint(f, yt, t, method="dopri5", rtol=1e-5, atol=1e-5, options=opts)
        dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            gode.odeint(f, yt, t, method="dopri5", rtol=1e-5, atol=1e-5, options=opts)
        e.record()
        torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3, gode.last_step_log().n_attempts


tw, na = timed({"norm": "world"})
tl, nl = timed({})
if rank == 0:
    res["fwd_us_B4096_per_rank"] = dict(world_norm=round(tw, 1), attempts_world=na, per_rank_norm=round(tl, 1), attempts_local=nl)
    os.write(out_fd, (json.dumps(res) + "\n").encode())
okt = torch.tensor([1 if ok else 0], device=dev)
dist.broadcast(okt, 0)
ok = bool(okt.item())
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
