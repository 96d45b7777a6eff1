#!/usr/bin/env python
"""bench.py — ODE trajectory-steps/s (fwd+bwd) of the latent-motion hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): the reference's ODEFunc
(models/mocogan_ode.py:6-17, D = H = 16, default nn.Linear init under torch.manual_seed(0)), B = 4096 trajectories
PER GPU (weak scaling), output grid t = linspace(0,1,16), dopri5 with rtol = atol = 1e-5, forward + backprop through
the solver, upstream gradient supplied as a resident N(0,1) tensor.  One "step" = one forward + one
backward of that batch.  trajectory-steps = B x ATTEMPTED dopri5 steps (accepted + rejected), read from the device log.

value     inputs resident in HBM; the step (2 kernels; the grid-sync workspace is persistent, so there are no memsets)
          replayed from CUDA graphs, K steps back to back over 20 distinct resident batches (together larger than L2, so
          no flush kernel is needed); the graphs hold --steps-per-graph (10) consecutive steps each, as a training loop
          that replays its inner loop from a graph would (inside a graph the kernel -> kernel gap is ~2 us, between two
          graph launches ~6 us); one CUDA-event pair around the K steps.  Beside it: run.one_step_per_graph_ms_per_step
          (the same K steps from single-step graphs, the round-1 definition) and run.latency_ms_per_step (isolated steps:
          L2 flush + event pair per step).
e2e       the same metric through the public replay API (gan_ode_b200.GraphedSolvePipeline, two slots in flight, five steps
          per slot / graph launch; e2e.one_step_per_slot_ms_per_step beside it) with the
          batch's noise y0 in pinned HOST memory: H2D copy of y0, forward + backward and D2H read of the parameter
          gradients are inside the timed region, every step; e2e.serial_* is one step at a time (GraphedSolveStep run +
          sync), e2e.eager_api_* the plain eager odeint + autograd call the reference makes.
roofline  dominant kernel, algorithmic HBM bytes / CUDA-event duration vs MEASURED_PEAKS.json.
--impl reference   the CPU arm: the torchdiffeq-restatement oracle (the real torchdiffeq is neither vendored by the
          reference nor installable here) on all host threads, same config / metric / unit.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

B_PER_GPU, D, H, T = 4096, 16, 16, 16
RTOL = ATOL = 1e-5
METRIC = "ode_trajectory_steps_per_s_fwd_bwd"
UNIT = "trajectory-steps/s"
CONFIG = {
    "workload": "configs[1]: ODEFunc D=H=16, B=4096 trajectories per GPU, t=linspace(0,1,16), dopri5 rtol=atol=1e-5, "
                "fwd + backprop-through-solver",
    "B_per_gpu": B_PER_GPU, "D": D, "H": H, "T": T, "rtol": RTOL, "atol": ATOL, "method": "dopri5",
    "backward": "backprop-through-solver", "weights": "nn.Linear default init, torch.manual_seed(0)",
    "l2": "inputs larger than L2: the timed steps rotate over 20 distinct resident batches (~10 MB touched per step, "
          "200 MB > 126 MB L2), no flush kernel between steps",
}


def config_for(n_gpus):
    """The `config` object both arms print: identical keys and values for the same --gpus N."""
    return dict(CONFIG, global_batch=B_PER_GPU * n_gpus, parallelism="dp{}".format(n_gpus))


def make_inputs(seed=0, device="cpu"):
    from gan_ode_b200.fields import make_field   # the package's mirror of the reference modules (no oracle import here)
    f = make_field(D, H, seed=0)
    g = torch.Generator().manual_seed(1000 + seed)
    y0 = torch.randn(B_PER_GPU, D, generator=g)
    grad = torch.randn(T, B_PER_GPU, D, generator=g)
    t = torch.linspace(0, 1, T).float()
    return f.to(device), y0.to(device), grad.to(device), t


# ---------------------------------------------------------------------------------------------------------------
def cpu_arm_time(steps, warmup, threads):
    """One step = the full workload (B=4096, dopri5 fwd + backprop) through the CPU oracle."""
    from oracle import torchdiffeq_restatement as tdq
    torch.set_num_threads(threads)
    f, y0, grad, t = make_inputs()
    params = list(f.parameters())

    def step():
        y = y0.clone().requires_grad_(True)
        sol = tdq.odeint(f, y, t, method="dopri5", rtol=RTOL, atol=ATOL)
        torch.autograd.grad(sol, [y] + params, grad)
        return len(tdq.last_step_log().accepted)

    for _ in range(warmup):
        n_att = step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        n_att = step()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    med = ts[len(ts) // 2]
    return B_PER_GPU * n_att / med, med, n_att


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps, warmup = max(1, args.steps), max(3, args.warmup)     # the same counts the GPU arm prints for these flags
    val, sec, n_att = cpu_arm_time(steps, warmup, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_for(args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "full workload per step (B=4096, {} attempted dopri5 steps), median of {} steps; "
                                   "torchdiffeq-restatement oracle (real torchdiffeq not installable)".format(n_att, steps)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy bandwidth)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measure_extras(gode, dev):
    """Side measurements (not the headline): where the kernels sit against their rooflines once the batch is large
    enough to leave the latency regime.  CUDA events, GPU kept busy while the host enqueues, median of 5."""
    from gan_ode_b200.fields import make_field

    def clone_to(f, device):
        return f.to(device)

    def timeit(fn, n=5, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for a, b in evs:
            torch.cuda._sleep(1000000)
            a.record(); fn(); b.record()
        torch.cuda.synchronize()
        return sorted(a.elapsed_time(b) for a, b in evs)[n // 2] * 1e-3  # s

    out = {}
    hbm, _ = peaks()
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    bf16 = float(json.load(open(pk)).get("bf16_tflops_sustained", 1348.6)) if os.path.exists(pk) else 1400.0
    t = torch.linspace(0, 1, 16).float()
    f16 = clone_to(make_field(16, 16, seed=0), dev)
    B = 1 << 20
    y0 = torch.randn(B, 16, device=dev)
    with torch.no_grad():
        for prec in ("fp32", "tf32", "bf16"):
            sec = timeit(lambda: gode.odeint(f16, y0, t, method="rk4", options={"precision": prec}))
            out["rk4_fwd_D16_H16_B1M_" + prec] = {
                "trajectory_steps_per_s": B * 15 / sec, "ms": sec * 1e3,
                "hbm_frac": (B * 15 * 64 / sec / 1e9) / hbm,      # 4*D bytes per trajectory-step (trajectory write)
                "fp32_ffma_tflops": (B * 15 * 4096 / sec / 1e12) if prec == "fp32" else None}
    B2 = 1 << 18
    y0r = torch.randn(B2, 16, device=dev, requires_grad=True)
    g = torch.randn(16, B2, 16, device=dev)
    sol = gode.odeint_adjoint(f16, y0r, t, method="rk4")
    sec_b = timeit(lambda: torch.autograd.grad(sol, [y0r] + list(f16.parameters()), g, retain_graph=True))
    with torch.no_grad():
        sec_f = timeit(lambda: gode.odeint(f16, y0r, t, method="rk4"))
    out["rk4_fwd_adjoint_D16_H16_B262144_fp32"] = {
        "trajectory_steps_per_s": B2 * 15 / (sec_f + sec_b), "fwd_ms": sec_f * 1e3, "bwd_ms": sec_b * 1e3,
        "fp32_tflops": B2 * 15 * 16384 / (sec_f + sec_b) / 1e12,
        "hbm_frac": (B2 * 15 * 192 / (sec_f + sec_b) / 1e9) / hbm}
    # the same shape with both directions on tcgen05 (bf16 operands): forward tc_rk4_fwd_kernel + adjoint tc_rk4_adj_small_kernel
    o16 = {"precision": "bf16", "bwd_precision": "bf16"}
    sol = gode.odeint_adjoint(f16, y0r, t, method="rk4", options=o16)
    sec_b = timeit(lambda: torch.autograd.grad(sol, [y0r] + list(f16.parameters()), g, retain_graph=True))
    with torch.no_grad():
        sec_f = timeit(lambda: gode.odeint(f16, y0r, t, method="rk4", options=o16))
    out["tc_rk4_fwd_adjoint_D16_H16_B262144_bf16"] = {
        "trajectory_steps_per_s": B2 * 15 / (sec_f + sec_b), "fwd_ms": sec_f * 1e3, "bwd_ms": sec_b * 1e3,
        "hbm_frac": (B2 * 15 * 192 / (sec_f + sec_b) / 1e9) / hbm}
    del y0, y0r, g, sol
    fw = clone_to(make_field(64, 256, seed=0), dev)
    for B3 in (148 * 128, 4 * 296 * 128):  # one tile per SM (latency regime) / four rounds of two tiles per SM
        yw = torch.randn(B3, 64, device=dev)
        with torch.no_grad():
            sec = timeit(lambda: gode.odeint(fw, yw, t, method="rk4", options={"precision": "bf16"}))
        tfl = B3 * 15 * 262144 / sec / 1e12
        # binding roofline of this shape is the MUFU pipe: one tanh per 4*D = 256 FLOP at 16 tanh/clk/SM (measured,
        # scripts/tmem_bench.cu) = half the tcgen05 rate at the same clock.  The bound is quoted at the MAXIMUM SM clock
        # (the kernel's actual clock under load is lower), so the fraction is a lower bound.
        mufu_tfl = 16 * 148 * 1965e6 * 256 / 1e12
        out["tc_rk4_fwd_D64_H256_B%d_bf16" % B3] = {"trajectory_steps_per_s": B3 * 15 / sec, "ms": sec * 1e3, "tflops": tfl,
                                                    "frac_of_measured_bf16_gemm": tfl / bf16,
                                                    "mufu_bound_tflops_at_1965mhz": mufu_tfl,
                                                    "frac_of_mufu_bound": tfl / mufu_tfl}
        del yw
    # forward + tensor-core continuous adjoint of the wide field (configs[3]'s shape), 64*D*H FLOP per trajectory-step
    B4 = 296 * 128
    yw = torch.randn(B4, 64, device=dev, requires_grad=True)
    gw = torch.randn(16, B4, 64, device=dev)
    solw = gode.odeint_adjoint(fw, yw, t, method="rk4", options={"precision": "bf16"})
    sec_b = timeit(lambda: torch.autograd.grad(solw, [yw] + list(fw.parameters()), gw, retain_graph=True))
    with torch.no_grad():
        sec_f = timeit(lambda: gode.odeint(fw, yw, t, method="rk4", options={"precision": "bf16"}))
    tfl = B4 * 15 * 1048576 / (sec_f + sec_b) / 1e12
    out["tc_rk4_fwd_adjoint_D64_H256_B%d_bf16" % B4] = {
        "trajectory_steps_per_s": B4 * 15 / (sec_f + sec_b), "fwd_ms": sec_f * 1e3, "bwd_ms": sec_b * 1e3, "tflops": tfl,
        "frac_of_measured_bf16_gemm": tfl / bf16}
    del yw, gw, solw
    # configs[2]: the ODE-RNN sampler, 16 frames of [dopri5 solve over [0,1] at torchdiffeq's default tolerances -> GRU jump],
    # B = 8192: one fused call per direction (gradient of the recorded steps) and the same with torchdiffeq's continuous
    # adjoint re-solve per frame (what the reference loop computes; also what it gets through the import shim)
    rnn = torch.nn.GRUCell(16, 16).to(dev)
    Br, Fr = 8192, 16
    h0 = torch.randn(Br, 16, device=dev, requires_grad=True)
    eps = torch.randn(Fr, Br, 16, device=dev)
    wgt = torch.randn(Fr, Br, 16, device=dev)
    for mode in ("discrete", "continuous"):
        def rnn_step():
            codes = gode.odernn_codes(f16, rnn, h0, eps, options={"adjoint": mode})
            torch.autograd.grad((codes * wgt).sum(), [h0] + list(f16.parameters()) + list(rnn.parameters()))
        sec = timeit(rnn_step)
        att = sum(fr["n_attempts"] for fr in gode.odernn.last_log().frames())
        out["odernn_fwd_bwd_B8192_F16_%s_adjoint" % mode] = {"ms": sec * 1e3, "attempted_steps_fwd": att,
                                                             "trajectory_steps_per_s": Br * att / sec}
    # configs[3] end to end: the reference's training loop (ucf_moco_ode.py:113-163) on synthetic clips, 32
    # videos, D=64 / H=256 motion ODE through the shim (scripts/train_dp_harness.py; reference nets when mounted)
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("train_dp_harness", os.path.join(ROOT, "scripts", "train_dp_harness.py"))
        hmod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(hmod)
        prev = (gode.config.layout, gode.config.precision)
        try:
            res = hmod.run_harness(iters=5, batch=32, precision="bf16", warmup=2)
        finally:
            gode.config.layout, gode.config.precision = prev
            sys.modules.pop("torchdiffeq", None)
            sys.modules.pop("torchsde", None)
        out["train_dp_harness_configs3"] = {k: res[k] for k in ("nets", "videos_per_gpu", "iters_per_s", "ms_per_iter",
                                                                "ode_forward_ms_per_iter", "ode_forward_share",
                                                                "ode_trajectories_per_iter")}
    except Exception as e:  # noqa: BLE001
        out["train_dp_harness_configs3"] = {"error": str(e)[:200]}
    return out


def measure_strong_scaling(gode, dev, rank, world):
    """N > 1 side measurements, run by EVERY rank: BASELINE.json configs[2] (ODE-RNN sampler, B = 8192 TOTAL, 16 frames,
    torchdiffeq default tolerances) and configs[4] (Euler-Maruyama, B = 16384 TOTAL, 41 steps, Philox keyed by the GLOBAL
    trajectory index) sharded over the ranks — fixed total work, forward + backward, parameter gradients all-reduced.
    CUDA events per step, barrier + synchronize on both sides, MAX over ranks."""
    import torch.distributed as dist
    from gan_ode_b200.dist import shard_bounds
    from gan_ode_b200.fields import ODEFunc, SDEFunc
    steps, warm = 10, 3
    prev, prev_x = gode.config.grad_allreduce, gode.config.grad_exchange
    if prev is None:
        gode.config.grad_allreduce = True

    def timed(fn):
        for _ in range(warm):
            fn()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        dist.barrier()
        torch.cuda.synchronize()
        for a, b in evs:
            a.record(); fn(); b.record()
        torch.cuda.synchronize()
        dist.barrier()
        tt = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item()) / steps

    out = {}
    try:
        torch.manual_seed(0)
        ode_fn, gru = ODEFunc(16, 16).to(dev), torch.nn.GRUCell(16, 16).to(dev)
        NB = 8192
        lo, hi = shard_bounds(NB, rank, world)
        g = torch.Generator().manual_seed(1)
        h0 = torch.randn(NB, 16, generator=g)[lo:hi].to(dev)
        eps = torch.randn(16, NB, 16, generator=g)[:, lo:hi].contiguous().to(dev)
        w = torch.randn(16, NB, 16, generator=g)[:, lo:hi].contiguous().to(dev)
        params = list(ode_fn.parameters()) + list(gru.parameters())

        def rnn_step():
            return torch.autograd.grad(gode.odernn_codes(ode_fn, gru, h0, eps), params, w)

        ms = timed(rnn_step)
        att = torch.tensor([sum(f["n_attempts"] for f in gode.odernn.last_log().frames())], device=dev)
        dist.all_reduce(att, op=dist.ReduceOp.MAX)
        out["configs2_odernn_B8192_total_16frames_default_tol_fwd_bwd"] = {
            "scaling": "strong", "ms_per_step": ms, "attempted_steps_all_frames": int(att.item()),
            "trajectory_steps_per_s": NB * int(att.item()) / ms * 1e3}

        sde = SDEFunc(16, 16).to(dev)
        NS = 16384
        lo, hi = shard_bounds(NS, rank, world)
        g = torch.Generator().manual_seed(2)
        y0 = torch.randn(NS, 16, generator=g)[lo:hi].to(dev).requires_grad_(True)
        gr = torch.randn(16, NS, 16, generator=g)[:, lo:hi].contiguous().to(dev)
        ts = torch.linspace(0, 1, 16).float()
        sp = list(sde.parameters())

        def sde_step():
            sol = gode.sdeint(sde, y0, ts, method="euler", dt=2.5e-2, bm=gode.PhiloxBrownian(1234, lo))
            return torch.autograd.grad(sol, [y0] + sp, gr)

        ms = timed(sde_step)
        out["configs4_sde_em_B16384_total_41steps_philox_fwd_bwd"] = {
            "scaling": "strong", "ms_per_step": ms, "trajectory_steps_per_s": NS * 41 / ms * 1e3}
    except Exception as e:  # noqa: BLE001
        out["error"] = str(e)[:200]
    gode.config.grad_allreduce, gode.config.grad_exchange = prev, prev_x
    return out


def run_gpu(args):
    import torch.distributed as dist
    import gan_ode_b200 as gode

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback on the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    out_fd = 1
    xchg_fallback = ""
    if world > 1:
        # stdout must carry exactly ONE JSON line (rank 0).  NCCL prints its version banner (and anything NCCL_DEBUG asks
        # for) on fd 1 from native code, so fd 1 is pointed at stderr for the whole run and the JSON line is written to
        # the saved descriptor at the end.
        sys.stdout.flush()
        out_fd = os.dup(1)
        os.dup2(2, 1)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
        from gan_ode_b200.dist import enable_fused_grad_exchange, enable_p2p_allreduce
        # parameter-gradient exchange, in order of preference: fused into the backward kernel's reduction tail over NVLink
        # peer memory; one-shot peer-memory kernel behind the backward; ncclAllReduce
        fused = args.grad_exchange == "fused" and enable_fused_grad_exchange(require=args.require_grad_exchange)
        if not fused:
            xchg_fallback = getattr(fused, "why", "")
            p2p = args.grad_exchange != "nccl" and enable_p2p_allreduce(
                require=args.require_grad_exchange and args.grad_exchange == "p2p")
            if not p2p:
                xchg_fallback = (xchg_fallback + " | " + getattr(p2p, "why", "")).strip(" |")
                gode.config.grad_allreduce = True
    n_gpus = world

    f, y0, grad, t = make_inputs(seed=rank, device=dev)
    params = list(f.parameters())
    kw = dict(method="dopri5", rtol=RTOL, atol=ATOL)
    y0r = y0.clone().requires_grad_(True)

    def step():
        sol = gode.odeint(f, y0r, t, **kw)
        return torch.autograd.grad(sol, [y0r] + params, grad)

    # warm-up (eager), read the step count from the device log
    for _ in range(max(3, args.warmup)):
        grads = step()
    torch.cuda.synchronize()
    log = gode.last_step_log()
    n_att, n_acc, status = log.n_attempts, log.n_accepted, log.status
    assert status == 0, "solver status {}".format(status)

    # capture fwd+bwd (+ the NCCL all-reduce when N>1) in a CUDA graph; fall back to eager launches if refused
    graph, graphed = None, False
    try:
        if args.no_graph:
            raise RuntimeError("--no-graph")
        # (--pdl: the backward captured as a programmatic dependent launch behind the forward, gan_ode_b200.config.pdl)
        gode.config.pdl = args.pdl
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_grads = step()
        gode.config.pdl = False
        graph.replay()
        torch.cuda.synchronize()
        ok = all(torch.allclose(a, b, rtol=1e-4, atol=1e-6) for a, b in zip(static_grads, grads))
        graphed = bool(ok)
        if not ok:
            graph = None
    except Exception as e:  # noqa: BLE001
        sys.stderr.write("[bench] CUDA-graph capture unavailable ({}); timing eager launches\n".format(str(e)[:200]))
        graph = None
        gode.config.pdl = False
        torch.cuda.synchronize()

    run_step = graph.replay if graph is not None else step
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def timed_region(fn, k):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        for a, b in evs:
            flush.fill_(1)
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        tot = sum(a.elapsed_time(b) for a, b in evs)  # ms
        tt = torch.tensor([tot], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    # ---- value: K steps back to back over distinct resident batches -------------------------------------------------------
    # One captured graph per input batch; N_SETS batches touch ~10 MB each, together more than the 126 MB L2, so every step
    # reads cold inputs WITHOUT a flush kernel in between and the steps pipeline the way a training loop issues them (the
    # next graph launch is in flight while the current one runs).  One CUDA-event pair brackets all K steps.  The isolated
    # per-step time (flush before every step, event pair per step: includes one graph-launch latency per step) is reported
    # beside it as `latency_ms_per_step`.
    N_SETS = 20
    sets = []     # (graph, n_attempts)
    if graph is not None:
        try:
            gode.config.pdl = args.pdl
            for i in range(N_SETS):
                gi = torch.Generator().manual_seed(1000 + rank + 7919 * (i + 1))
                yi = torch.randn(B_PER_GPU, D, generator=gi).to(dev).requires_grad_(True)
                gri = torch.randn(T, B_PER_GPU, D, generator=gi).to(dev)

                def step_i(yi=yi, gri=gri):
                    sol = gode.odeint(f, yi, t, **kw)
                    return torch.autograd.grad(sol, [yi] + params, gri)

                s_ = torch.cuda.Stream()
                s_.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s_):
                    step_i()
                torch.cuda.current_stream().wait_stream(s_)
                torch.cuda.synchronize()
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_):
                    keep_i = step_i()
                log_i = gode.last_step_log()
                g_.replay()
                torch.cuda.synchronize()
                assert log_i.status == 0
                sets.append((g_, int(log_i.n_attempts), keep_i, yi, gri))
        except Exception as e:  # noqa: BLE001
            sys.stderr.write("[bench] batch rotation unavailable ({}); timing isolated steps\n".format(str(e)[:200]))
            sets = []
        gode.config.pdl = False

    # The same steps captured SEVERAL PER GRAPH (a training loop that replays its inner loop from a graph does this): between two
    # kernels of one graph the launch gap is ~2 us, between two graph launches on a stream ~6 us (scripts/
    # multistep_graph_probe.py: 51.0 / 47.8 / 47.2 / 46.1 us per step at 1 / 2 / 4 / 8 steps per graph).  Group g replays the
    # steps of batches g*S .. g*S+S-1; a K that is not a multiple of S finishes on the single-step graphs.
    STEPS_PER_GRAPH = max(1, args.steps_per_graph)
    groups = []   # (graph, attempted steps of its S batches)
    if sets and STEPS_PER_GRAPH > 1:
        try:
            gode.config.pdl = args.pdl
            for g0 in range(0, len(sets) - STEPS_PER_GRAPH + 1, STEPS_PER_GRAPH):
                members = sets[g0:g0 + STEPS_PER_GRAPH]
                s_ = torch.cuda.Stream()
                s_.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s_):
                    torch.autograd.grad(gode.odeint(f, members[0][3], t, **kw), [members[0][3]] + params, members[0][4])
                torch.cuda.current_stream().wait_stream(s_)
                torch.cuda.synchronize()
                gg = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gg):
                    keep_g = [torch.autograd.grad(gode.odeint(f, m[3], t, **kw), [m[3]] + params, m[4]) for m in members]
                gg.replay()
                torch.cuda.synchronize()
                assert all(torch.equal(a_, b_) for kg, m in zip(keep_g, members) for a_, b_ in zip(kg, m[2])), \
                    "multi-step graph disagrees with the single-step graphs"
                groups.append((gg, sum(m[1] for m in members), keep_g))
        except Exception as e:  # noqa: BLE001
            sys.stderr.write("[bench] multi-step graphs unavailable ({}); one step per graph\n".format(str(e)[:200]))
            groups = []
        gode.config.pdl = False

    def throughput_region(k, grouped=True):
        """K steps over the rotating batches, one event pair; returns (max-over-ranks ms, trajectory-steps of ALL ranks)."""
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        use = groups if grouped else []
        n_group_launches = (k // STEPS_PER_GRAPH) if use else 0
        rest = k - n_group_launches * STEPS_PER_GRAPH
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a.record()
        for j in range(n_group_launches):
            use[j % len(use)][0].replay()
        for j in range(rest):
            sets[j % len(sets)][0].replay()
        b.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        units = float(B_PER_GPU * (sum(use[j % len(use)][1] for j in range(n_group_launches)) +
                                   sum(sets[j % len(sets)][1] for j in range(rest))))
        tt = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
        uu = torch.tensor([units], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(uu, op=dist.ReduceOp.SUM)
        return float(tt.item()), float(uu.item())

    for _ in range(max(3, args.warmup)):
        run_step()
    sampler = ClockSampler(local) if rank == 0 else None
    latency_ms = timed_region(run_step, args.steps)
    single_ms = None
    if sets:
        throughput_region(max(3, args.warmup) * STEPS_PER_GRAPH)
        total_ms, total_units = throughput_region(args.steps)
        if groups:
            single_ms, _ = throughput_region(args.steps, grouped=False)
    else:
        total_ms, total_units = latency_ms, None
    eager_ms = timed_region(step, args.steps) if graph is not None else latency_ms

    # N > 1: name what the step costs beyond N = 1.  The same K replays over the same batches with the gradient exchange
    # taken out (every rank runs alone): the slowest rank's own rate is what lock-step exchange can reach at best (rank
    # skew); the rest of the difference to the timed value is the exchange itself (NVLink store -> poll latency).
    skew = None
    if world > 1 and sets and not args.no_extras:
        try:
            prev, prev_x = gode.config.grad_allreduce, gode.config.grad_exchange
            gode.config.grad_allreduce = gode.config.grad_exchange = None
            alone = []
            for (_, na_, _, yi, gri) in sets:
                def step_a(yi=yi, gri=gri):
                    return torch.autograd.grad(gode.odeint(f, yi, t, **kw), [yi] + params, gri)
                s_ = torch.cuda.Stream()
                s_.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s_):
                    step_a()
                torch.cuda.current_stream().wait_stream(s_)
                torch.cuda.synchronize()
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_):
                    keep_a = step_a()
                alone.append((g_, keep_a))
            gode.config.grad_allreduce, gode.config.grad_exchange = prev, prev_x
            for rep in range(2):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                dist.barrier()
                torch.cuda.synchronize()
                a.record()
                for j in range(args.steps):
                    alone[j % len(alone)][0].replay()
                b.record()
                torch.cuda.synchronize()
            with_x = single_ms if single_ms is not None else total_ms    # like for like: one step per graph on both sides
            mine = torch.tensor([a.elapsed_time(b) / args.steps], device=dev, dtype=torch.float64)
            every = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(every, mine)
            per_rank = [float(x.item()) for x in every]
            skew = {"what": "same {} single-step graph replays over the same batches WITHOUT the gradient exchange, every rank "
                            "alone; ms per step".format(args.steps),
                    "no_exchange_ms_per_step_per_rank": per_rank,
                    "no_exchange_slowest_rank_ms": max(per_rank), "no_exchange_fastest_rank_ms": min(per_rank),
                    "with_exchange_ms_per_step": with_x / args.steps,
                    "rank_skew_us": (max(per_rank) - min(per_rank)) * 1e3,
                    "exchange_cost_over_slowest_rank_us": (with_x / args.steps - max(per_rank)) * 1e3}
            del alone
        except Exception as e:  # noqa: BLE001
            skew = {"error": str(e)[:200]}
            gode.config.grad_allreduce, gode.config.grad_exchange = prev, prev_x

    # ---- e2e through the public API with host buffers -----------------------------------------------------------
    # (a) gan_ode_b200.GraphedSolveStep: the public replay API — H2D(y0 pinned) + fwd + bwd + D2H(param grads pinned)
    #     in one graph launch per step, then a stream sync so the host can read the result.
    # (b) the plain eager call the reference makes (odeint + autograd), for comparison.
    def e2e_time(fn, k):
        for _ in range(3):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(k):
            fn()
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        te = torch.tensor([sec], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return float(te.item())

    y0_host = y0.cpu().pin_memory()
    n_param = sum(p.numel() for p in params)
    e2e_serial_s = None
    e2e_one_s = None
    e2e_api = "gan_ode_b200.GraphedSolveStep (H2D y0 + odeint fwd + backprop + D2H param grads, one graph launch, sync)"
    try:
        if args.no_graph:
            raise RuntimeError("--no-graph")
        gs = gode.GraphedSolveStep(f, B_PER_GPU, t, adjoint=False, read_back=("param_grads",), pdl=args.pdl, **kw)
        gs.y0_host.copy_(y0_host)
        gs.grad_traj.copy_(grad)

        def e2e_step():
            gs.run()
            return gs.sync()["param_grads"]

        chk = e2e_step().to(dev)
        ref_flat = torch.cat([g.reshape(-1) for g in grads[1:]])
        assert torch.allclose(chk, ref_flat, rtol=1e-4, atol=1e-6), "graphed e2e step disagrees with the eager step"
        e2e_s = e2e_time(e2e_step, args.steps)
        e2e_serial_s = e2e_s
        # (a') the same step through GraphedSolvePipeline, two steps in flight: step i+1's H2D crosses PCIe while step i
        #      computes; every step still has its own H2D and D2H and its result is read on the host before the step
        #      after next is issued.
        if not args.no_pipeline:
            pipe = gode.GraphedSolvePipeline(f, B_PER_GPU, t, depth=2, adjoint=False, read_back=("param_grads",),
                                             pdl=args.pdl, **kw)
            for sl in pipe.slots:
                sl.y0_host.copy_(y0_host)
                sl.grad_traj.copy_(grad)
            pipe.submit()
            chk = pipe.result()["param_grads"].to(dev)
            assert torch.allclose(chk, ref_flat, rtol=1e-4, atol=1e-6), "pipelined e2e step disagrees with the eager step"

            # ... and with E2E_S steps per slot (one graph launch carries E2E_S steps, each with its own input batch in the
            # pinned block and its own gradient read back): the launch gap between graphs is amortised as in `value`
            E2E_S = max(1, args.e2e_steps_per_slot)
            pipe_s = None
            if E2E_S > 1:
                pipe_s = gode.GraphedSolvePipeline(f, B_PER_GPU, t, depth=2, adjoint=False, read_back=("param_grads",),
                                                   pdl=args.pdl, steps=E2E_S, **kw)
                for sl in pipe_s.slots:
                    for k_ in range(E2E_S):
                        sl.y0_host[k_].copy_(y0_host)
                        sl.grad_traj[k_].copy_(grad)
                pipe_s.submit()
                chk = pipe_s.result()["param_grads"].to(dev)
                assert all(torch.allclose(chk[k_], ref_flat, rtol=1e-4, atol=1e-6) for k_ in range(E2E_S)), \
                    "multi-step pipelined e2e disagrees with the eager step"

            def drive(pp, n_submit):
                n_read = 0
                for _ in range(n_submit):
                    if len(pp._inflight) == len(pp.slots):
                        pp.result()
                        n_read += 1
                    pp.submit()
                while pp._inflight:
                    pp.result()
                    n_read += 1
                assert n_read == n_submit

            def e2e_pipelined(k):
                n_groups = (k // E2E_S) if pipe_s is not None else 0
                if n_groups:
                    drive(pipe_s, n_groups)
                drive(pipe, k - n_groups * E2E_S)

            e2e_pipelined(3 * E2E_S + 2)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e2e_pipelined(args.steps)
            torch.cuda.synchronize()
            te = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            e2e_s = float(te.item())
            e2e_api = ("gan_ode_b200.GraphedSolvePipeline, 2 slots in flight, {} step(s) per slot (per step: H2D of its y0 from "
                       "pinned host, odeint fwd + backprop, D2H of its param grads, result read on the host; a slot's H2D "
                       "overlaps the previous slot's kernels; {} slot launches + {} single-step launches)".format(
                           E2E_S if pipe_s is not None else 1, (args.steps // E2E_S) if pipe_s is not None else 0,
                           args.steps - ((args.steps // E2E_S) * E2E_S if pipe_s is not None else 0)))
            # the single-step-per-slot figure beside it
            e2e_one_s = None
            if pipe_s is not None:
                drive(pipe, 6)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                drive(pipe, args.steps)
                torch.cuda.synchronize()
                te = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(te, op=dist.ReduceOp.MAX)
                e2e_one_s = float(te.item())
    except Exception as e:  # noqa: BLE001
        sys.stderr.write("[bench] GraphedSolveStep unavailable ({}); e2e falls back to the eager API\n".format(str(e)[:200]))
        e2e_s = None

    res_host = torch.empty(n_param, dtype=torch.float32).pin_memory()

    def e2e_eager_step():
        y = y0_host.to(dev, non_blocking=True).requires_grad_(True)
        sol = gode.odeint(f, y, t, **kw)
        gs_ = torch.autograd.grad(sol, params, grad)
        res_host.copy_(torch.cat([g.reshape(-1) for g in gs_]), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return res_host

    e2e_eager_s = e2e_time(e2e_eager_step, args.steps)
    # the same call with autograd's worker threads off (torch.autograd.set_multithreading_enabled(False)): the CUDA backward
    # node then runs on the calling thread instead of being handed to the engine's device thread and back
    e2e_eager_st_s = None
    try:
        with torch.autograd.set_multithreading_enabled(False):
            e2e_eager_st_s = e2e_time(e2e_eager_step, args.steps)
    except Exception:  # noqa: BLE001
        pass
    if e2e_s is None:
        e2e_s, e2e_api = e2e_eager_s, "gan_ode_b200.odeint (eager, pinned host y0)"
    clocks = sampler.stop() if sampler else None

    # ---- N>1: the fused NVLink gradient exchange against ncclAllReduce of the local gradients (outside the timed region)
    parity = None
    if world > 1:
        prev, prev_x = gode.config.grad_allreduce, gode.config.grad_exchange
        exchanged = [g.clone() for g in step()[1:]]                      # the path the timed region ran
        gode.config.grad_allreduce = gode.config.grad_exchange = None
        local_g = [g.clone() for g in step()[1:]]                        # this rank's own sums, no exchange
        gode.config.grad_allreduce, gode.config.grad_exchange = prev, prev_x
        flat_x = torch.cat([g.reshape(-1) for g in exchanged])
        flat_n = torch.cat([g.reshape(-1) for g in local_g])
        dist.all_reduce(flat_n, op=dist.ReduceOp.SUM)                    # NCCL, the library baseline
        gathered = [torch.empty_like(flat_x) for _ in range(world)]
        dist.all_gather(gathered, flat_x)
        max_rel = float((flat_x - flat_n).abs().max() / flat_n.abs().max())
        parity = {"what": "parameter gradients of the timed step's exchange path vs ncclAllReduce(sum) of the per-rank gradients",
                  "max_rel": max_rel, "ranks_identical": bool(all(torch.equal(x, gathered[0]) for x in gathered)),
                  "n_floats": int(flat_x.numel()), "ok": bool(max_rel <= 1e-5)}

    # ---- per-kernel durations for the roofline ---------------------------------------------------------------------------
    # Taken from graph replays so that they add up to the step: a graph holding only the forward launch is timed exactly
    # like the step graph (per-replay CUDA events, L2 flushed in between); the backward's share is the difference.  Eager
    # event brackets (round 1) also caught the launch gaps and over-stated both.
    roof = None
    if rank == 0:
        k = max(10, min(args.steps, 50))
        prev, prev_x = gode.config.grad_allreduce, gode.config.grad_exchange   # rank 0 alone: no exchange in here
        gode.config.grad_allreduce = gode.config.grad_exchange = None

        def fwd_only():
            return gode.odeint(f, y0r, t, **kw)

        def graph_of(fn):
            gode.config.pdl = args.pdl
            s_ = torch.cuda.Stream()
            s_.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s_):
                for _ in range(2):
                    fn()
            torch.cuda.current_stream().wait_stream(s_)
            torch.cuda.synchronize()
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_):
                keep_ = fn()
            gode.config.pdl = False
            return g_, keep_

        def replay_us(g_):
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
            for _ in range(3):
                g_.replay()
            for a, b in evs:
                flush.fill_(1)
                a.record(); g_.replay(); b.record()
            torch.cuda.synchronize()
            return sorted(a.elapsed_time(b) for a, b in evs)[k // 2] * 1e3

        timing_how = "CUDA-graph replays: forward-only graph; backward = (forward+backward graph) - forward"
        try:
            if args.no_graph:
                raise RuntimeError("--no-graph")
            gf, _kf = graph_of(fwd_only)
            gs1, _ks = graph_of(step)
            fwd_us = replay_us(gf)
            step_us = replay_us(gs1)
            bwd_us = max(step_us - fwd_us, 0.0)
            del gf, gs1
        except Exception as e:  # noqa: BLE001
            sys.stderr.write("[bench] per-kernel graph timing unavailable ({}); eager event brackets\n".format(str(e)[:200]))
            timing_how = "eager CUDA-event brackets (includes launch gaps)"
            ef = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(k)]
            for e0, e1, e2 in ef:
                flush.fill_(1)
                torch.cuda._sleep(400000)
                e0.record()
                sol = gode.odeint(f, y0r, t, **kw)
                e1.record()
                torch.autograd.grad(sol, [y0r] + params, grad)
                e2.record()
            torch.cuda.synchronize()
            fwd_us = sorted(e0.elapsed_time(e1) for e0, e1, _ in ef)[k // 2] * 1e3
            bwd_us = sorted(e1.elapsed_time(e2) for _, e1, e2 in ef)[k // 2] * 1e3
        gode.config.grad_allreduce, gode.config.grad_exchange = prev, prev_x
        row = B_PER_GPU * D * 4
        fwd_bytes = row * (1 + T + n_acc)           # read y0, write T outputs, write n_acc checkpoints
        bwd_bytes = row * (T + n_acc + 1) + 4 * n_param  # read T upstream grads + n_acc checkpoints, write grad_y0 + params
        dom, dur, byts = ("dopri5_backprop_bwd_kernel", bwd_us, bwd_bytes) if bwd_us >= fwd_us else \
                         ("dopri5_fwd_kernel", fwd_us, fwd_bytes)
        peak, how = peaks()
        achieved = byts / (dur * 1e-6) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(dom)
        roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": how,
                "algorithmic_bytes_per_launch": byts, "kernel_us": dur,
                "kernels_us": {"dopri5_fwd_kernel": fwd_us, "dopri5_backprop_bwd_kernel": bwd_us},
                "kernel_timing": timing_how,
                "step_bytes": fwd_bytes + bwd_bytes,
                "step_frac": (fwd_bytes + bwd_bytes) / (total_ms / args.steps * 1e-3) / 1e9 / peak,
                "note": "B=4096 x D=16 is 256 KB of state: the solve is a chain of {} dependent stages with a grid "
                        "barrier per attempted step, i.e. latency-bound, not bandwidth-bound".format(2 + 6 * n_att)}

    def finish():
        # Captured graphs hold NCCL kernels: release them before the communicator goes away, and do not call
        # destroy_process_group() (it can wait forever on a communicator that graphs still reference).
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            os._exit(0)

    strong = None
    if world > 1 and not args.no_extras:
        strong = measure_strong_scaling(gode, dev, rank, world)
    if rank != 0:
        finish()
        return

    units = B_PER_GPU * n_att * n_gpus            # per step, the batch of set 0 (latency / e2e figures)
    ms_per_step = total_ms / args.steps
    value = (total_units / (total_ms * 1e-3)) if total_units is not None else units / (ms_per_step * 1e-3)
    e2e_val = units / (e2e_s / args.steps)
    h2d = y0_host.numel() * 4
    d2h = n_param * 4

    # CPU baseline beside it: the oracle on the host cores, bounded sample (rank 0, N=1 only)
    cpu = None
    if n_gpus == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, sec, na = cpu_arm_time(steps=7, warmup=2, threads=threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "full workload (B=4096, {} attempted steps) x 7 timed steps, median {:.3f} s/step; "
                         "torchdiffeq-restatement oracle (PyTorch CPU)".format(na, sec)}

    extras = {"strong_scaling": strong} if strong is not None else None
    if n_gpus == 1 and not args.no_extras:
        try:
            extras = measure_extras(gode, dev)
        except Exception as e:  # noqa: BLE001
            extras = {"error": str(e)[:200]}

    # second structured roofline entry: the contraction-bound shape (configs[3]: D=64, H=256, rk4 forward + tensor-core
    # continuous adjoint), 64*D*H algorithmic FLOP per trajectory-step against the measured sustained BF16 GEMM rate
    roof_tc = None
    if isinstance(extras, dict):
        for key, v in extras.items():
            if key.startswith("tc_rk4_fwd_adjoint_D64_H256") and isinstance(v, dict) and "tflops" in v:
                pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
                bf16 = float(json.load(open(pk)).get("bf16_tflops_sustained", 1348.6)) if os.path.exists(pk) else 1400.0
                roof_tc = {"bound": "tensor", "kernel": "tc_rk4_fwd_wide2_kernel + tc_rk4_adj_wide_kernel", "workload": key,
                           "achieved": v["tflops"], "peak": bf16, "unit": "TFLOP/s", "frac": v["tflops"] / bf16,
                           "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if os.path.exists(pk) else "fallback",
                           "algorithmic_flop_per_trajectory_step": 64 * 64 * 256, "traffic": None,
                           "fwd_ms": v["fwd_ms"], "bwd_ms": v["bwd_ms"],
                           "trajectory_steps_per_s": v["trajectory_steps_per_s"]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config_for(n_gpus),
        "run": {"cuda_graph": graphed, "attempted_steps": n_att, "accepted_steps": n_acc,
                "timing": ("{} steps replayed back to back over {} distinct resident batches (no flush kernel; one CUDA-event "
                           "pair around all steps); {}".format(
                               args.steps, len(sets),
                               "CUDA graphs of {} consecutive steps each ({} graph launches + {} single-step graphs)".format(
                                   STEPS_PER_GRAPH, args.steps // STEPS_PER_GRAPH, args.steps % STEPS_PER_GRAPH)
                               if groups else "one CUDA graph per step") if sets else
                           "isolated steps: L2 flush + one CUDA-event pair per step"),
                "steps_per_graph": STEPS_PER_GRAPH if groups else 1,
                "one_step_per_graph_ms_per_step": (single_ms / args.steps) if single_ms is not None else None,
                "attempted_steps_per_batch": [x[1] for x in sets] if sets else [n_att],
                "latency_ms_per_step": latency_ms / args.steps,
                "latency_value": units / (latency_ms / args.steps * 1e-3),
                "pdl_backward": bool(graphed and args.pdl),
                "grad_allreduce": ("none (1 GPU)" if n_gpus == 1 else
                                   "fused into the backward kernel's reduction tail over NVLink peer memory "
                                   "(gode_dopri5_backprop_bwd_world)" if gode.config.grad_exchange is not None else
                                   "one-shot kernel over NVLink peer memory (csrc/p2p_allreduce.cu)"
                                   if callable(gode.config.grad_allreduce) else "ncclAllReduce"),
                "multi_gpu_step_cost": skew,
                "grad_exchange_requested": args.grad_exchange if n_gpus > 1 else None,
                "grad_exchange_fallback_reason": xchg_fallback or None},
        "parity_check": parity,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s / args.steps * 1e3, "api": e2e_api,
                "one_step_per_slot_ms_per_step": (e2e_one_s / args.steps * 1e3) if e2e_one_s else None,
                "serial_ms_per_step": (e2e_serial_s / args.steps * 1e3) if e2e_serial_s else None,
                "serial_value": (units / (e2e_serial_s / args.steps)) if e2e_serial_s else None,
                "serial_api": "gan_ode_b200.GraphedSolveStep.run() + .sync() per step (no overlap between steps)",
                "eager_api_value": units / (e2e_eager_s / args.steps), "eager_api_ms_per_step": e2e_eager_s / args.steps * 1e3,
                "eager_api_autograd_single_thread_ms_per_step": (e2e_eager_st_s / args.steps * 1e3) if e2e_eager_st_s else None},
        "gpu_launches": (2 + (1 if n_gpus > 1 and gode.config.grad_exchange is None and callable(gode.config.grad_allreduce)
                              else 0)) * args.steps,
        "eager_ms_per_step": eager_ms / args.steps,
        "clocks": clocks, "roofline": roof, "roofline_configs3": roof_tc, "cpu_baseline": cpu, "extras": extras,
    }
    sys.stdout.flush()
    os.write(out_fd, (json.dumps(line) + "\n").encode())
    finish()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replay")
    ap.add_argument("--grad-exchange", choices=("fused", "p2p", "nccl"), default="fused",
                    help="N>1: how the parameter gradient is summed over ranks (default: inside the backward kernel)")
    ap.add_argument("--steps-per-graph", type=int, default=10, help="value: consecutive steps captured per CUDA graph "
                    "(1: one graph per step, the round-1 definition; reported beside the value either way)")
    ap.add_argument("--e2e-steps-per-slot", type=int, default=5, help="e2e: steps carried by one pipeline slot / graph launch "
                    "(each with its own H2D input and D2H result); 1: one step per launch")
    ap.add_argument("--no-pipeline", action="store_true", help="e2e: one step at a time (GraphedSolveStep run + sync) "
                    "instead of two steps in flight")
    ap.add_argument("--require-grad-exchange", action="store_true", help="N>1: fail instead of falling back when the "
                    "requested peer-memory exchange cannot be set up")
    ap.add_argument("--nccl-allreduce", action="store_true", help="N>1: all-reduce the parameter gradient with NCCL instead "
                    "of the fused peer-memory kernel")
    ap.add_argument("--pdl", action="store_true", help="capture the backward as a programmatic dependent launch behind the "
                    "forward (measured slower for this pair of kernels; off by default)")
    ap.add_argument("--no-extras", action="store_true", help="skip the large-batch / wide-field side measurements")
    args = ap.parse_args()
    if args.nccl_allreduce:
        args.grad_exchange = "nccl"
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__), "--gpus", str(args.gpus),
               "--steps", str(args.steps), "--warmup", str(args.warmup)]
        sys.exit(subprocess.call(cmd))
    run_gpu(args)


if __name__ == "__main__":
    main()
