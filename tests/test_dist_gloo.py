"""world_size-2 gloo tests of the data-parallel host logic (CPU): shard bounds, the in-place gradient all-reduce hook
the backward Functions call, and that sharded gradients sum to the full-batch gradient.  The per-shard gradient is
produced here by the CPU oracle (standing in for the kernel, which needs a GPU); the GPU path is exercised by
`bench.py --gpus N` under gpurun."""
import importlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        api = importlib.import_module("gan_ode_b200.odeint")
        from gan_ode_b200 import dist as gdist
        from oracle import torchdiffeq_restatement as tdq
        from tests.helpers import make_field

        f = make_field(seed=0)
        g = torch.Generator().manual_seed(1)
        B = 37
        y0 = torch.randn(B, 16, generator=g)
        up = torch.randn(16, B, 16, generator=g)
        t = torch.linspace(0, 1, 16)
        local, lo = gdist.shard_batch(y0)
        hi = lo + local.shape[0]
        yl = local.clone().requires_grad_(True)
        sol = tdq.odeint_adjoint(f, yl, t, method="rk4")
        grads = torch.autograd.grad((sol * up[:, lo:hi]).sum(), [yl] + list(f.parameters()))
        flat = torch.cat([x.reshape(-1) for x in grads[1:]])
        gdist.enable_grad_allreduce()
        api._maybe_allreduce(flat)            # what _split_params does with the kernel's flat buffer
        gdist.disable_grad_allreduce()
        before = flat.clone()
        api._maybe_allreduce(flat)            # off again: must be a no-op
        assert torch.equal(before, flat)
        if rank == 0:
            yf = y0.clone().requires_grad_(True)
            full = torch.autograd.grad((tdq.odeint_adjoint(f, yf, t, method="rk4") * up).sum(), [yf] + list(f.parameters()))
            ref = torch.cat([x.reshape(-1) for x in full[1:]])
            out.put(("ok", float((flat - ref).abs().max() / ref.abs().max()), lo, hi,
                     float((grads[0] - full[0][lo:hi]).abs().max())))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sharded_gradients_allreduce_to_full_batch_gradient():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    status, rel, lo, hi, gy = out.get(timeout=10)
    assert status == "ok" and (lo, hi) == (0, 19)
    assert rel <= 1e-6 and gy <= 1e-6  # all-reduced param grad equals the single-process sum (fp32 1e-6)


def test_shard_bounds_partition_the_batch():
    from gan_ode_b200.dist import shard_bounds
    for B in (1, 7, 16, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
