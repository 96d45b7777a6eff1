"""CUDA-graph replay of one solve step (forward + backward + the host<->device copies around it).

At the reference's shapes (D = H = 16, a few thousand trajectories) one fused solve is tens of microseconds of GPU
time, so the host-side cost of *issuing* it (Python, autograd bookkeeping, two launches, the copies)
dominates unless the whole step is replayed from a graph.  `GraphedSolveStep` captures, once:

    H2D   y0 (pinned host)  ->  device
    fwd   odeint / odeint_adjoint (same kernels, same arguments as the eager call)
    bwd   gradients w.r.t. y0 and the ODEFunc parameters for a given upstream gradient
    [NCCL all-reduce of the flat parameter gradient when config.grad_allreduce is set]
    D2H   flat parameter gradient (and optionally grad_y0 / the trajectory)  ->  pinned host

and `run()` replays it with a single cudaGraphLaunch.  Parameters are read through their device addresses at replay
time, so in-place optimiser updates are picked up; rebuilding is only needed if shapes or the time grid change.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

import importlib

_api = importlib.import_module(__package__ + ".odeint")  # the module, not the re-exported function

__all__ = ["GraphedSolveStep", "GraphedSolvePipeline"]


def _flat_param_grads(gs):
    """The flat [W1|b1|W2|b2] gradient.  The backward kernels write ONE flat buffer and hand autograd views of it; when the
    views are still those (same base, back to back) the base is returned as it is — no concatenation kernel."""
    base = getattr(gs[0], "_base", None)
    if base is not None and base.dim() == 1 and all(getattr(g, "_base", None) is base for g in gs):
        off = gs[0].storage_offset()
        ok = True
        for g in gs:
            ok = ok and g.is_contiguous() and g.storage_offset() == off
            off += g.numel()
        if ok:
            return base[gs[0].storage_offset() - base.storage_offset():off - base.storage_offset()]
    return torch.cat([g.reshape(-1) for g in gs])


class GraphedSolveStep:
    def __init__(self, func, batch: int, t: torch.Tensor, *, method: Optional[str] = None, rtol=1e-7, atol=1e-9,
                 adjoint: bool = True, options: Optional[dict] = None, device=None,
                 read_back: Sequence[str] = ("param_grads",), warmup: int = 3, pdl: bool = False,
                 copies_in_graph: bool = True, steps: int = 1):
        """steps = S > 1: ONE graph holds S consecutive steps, each with its own input batch, upstream gradient and results
        (buffers get a leading S dimension: y0_host (S,B,D), grad_traj (S,T,B,D), host[...] (S,...)).  Between two kernels of
        one graph the launch gap is ~2 us, between two graph launches ~6 us, so S steps per launch amortise it."""
        W1, _, _, _ = _api.recognise_field(func)
        D = W1.shape[1]
        dev = torch.device(device) if device is not None else W1.device
        assert dev.type == "cuda", "GraphedSolveStep needs the ODEFunc on a CUDA device"
        self.func, self.t, self.device = func, t, dev
        self.params = _api.field_parameters(func)
        self._solve = _api.odeint_adjoint if adjoint else _api.odeint
        self._kw = dict(method=method, rtol=rtol, atol=atol, options=options)
        T = len(t)
        n_param = sum(p.numel() for p in self.params)
        self.read_back = tuple(read_back)
        for name in self.read_back:
            if name not in ("param_grads", "grad_y0", "traj"):
                raise ValueError("read_back entries must be 'param_grads', 'grad_y0' or 'traj'")

        # static buffers: pinned host side and device side (a leading `steps` dimension, dropped for steps == 1)
        S = self.steps = int(steps)
        assert S >= 1

        def lead(x):
            return x[0] if S == 1 else x

        self._y0_host_all = torch.zeros(S, batch, D, dtype=torch.float32).pin_memory()
        self._y0_all = torch.zeros(S, batch, D, dtype=torch.float32, device=dev)      # one H2D fills every step's input
        self._y0_steps = [self._y0_all[k].detach().requires_grad_(True) for k in range(S)]   # leaves sharing that storage
        self._grad_traj_all = torch.zeros(S, T, batch, D, dtype=torch.float32, device=dev)
        self.y0_host, self.y0, self.grad_traj = lead(self._y0_host_all), self._y0_steps[0], lead(self._grad_traj_all)
        self._host_all = {}
        if "param_grads" in self.read_back:
            self._host_all["param_grads"] = torch.zeros(S, n_param, dtype=torch.float32).pin_memory()
        if "grad_y0" in self.read_back:
            self._host_all["grad_y0"] = torch.zeros(S, batch, D, dtype=torch.float32).pin_memory()
        if "traj" in self.read_back:
            self._host_all["traj"] = torch.zeros(S, T, batch, D, dtype=torch.float32).pin_memory()
        self.host = {k: lead(v) for k, v in self._host_all.items()}
        self.traj = None
        self.grads = None
        self.log = None
        self.copies_in_graph = bool(copies_in_graph)   # False: GraphedSolvePipeline issues the copies around the graph
        self._d2h = []                                  # (pinned host buffer, device source) pairs of the last capture

        # In this graph the kernel right before the backward IS the matching forward (nothing in between writes the weights
        # or the upstream gradient), so the backward MAY be a programmatic dependent launch (config.pdl, `pdl=True`).  Off by
        # default: measured slower for the dopri5 pair (the 256-CTA backward does not fit beside the forward, the CTAs that
        # start late arrive late at its final reduction; profiles/README.md).
        prev_pdl = _api.config.pdl
        _api.config.pdl = bool(pdl)
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):
                    self._body()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._body()
        finally:
            _api.config.pdl = prev_pdl
        self.stream = torch.cuda.current_stream(dev)

    def _body(self):
        if self.copies_in_graph:
            self.copy_in()
        self._d2h = []
        for k in range(self.steps):
            y0 = self._y0_steps[k]
            sol = self._solve(self.func, y0, self.t, **self._kw)
            self.log = _api.last_step_log() if (self._kw["method"] in (None, "dopri5")) else None
            grads = torch.autograd.grad(sol, [y0] + self.params, self._grad_traj_all[k])
            self.traj, self.grads = sol.detach(), grads      # (of the last step)
            if "param_grads" in self._host_all:
                self._d2h.append((self._host_all["param_grads"][k], _flat_param_grads(grads[1:])))
            if "grad_y0" in self._host_all:
                self._d2h.append((self._host_all["grad_y0"][k], grads[0]))
            if "traj" in self._host_all:
                self._d2h.append((self._host_all["traj"][k], sol.detach()))
        if self.copies_in_graph:
            self.copy_out()

    def copy_in(self):
        """H2D of the pinned input on the current stream (a graph node when captured)."""
        with torch.no_grad():
            self._y0_all.copy_(self._y0_host_all, non_blocking=True)

    def copy_out(self):
        """D2H of the requested results on the current stream (graph nodes when captured)."""
        for host, src in self._d2h:
            host.copy_(src, non_blocking=True)

    def run(self, y0_host: Optional[torch.Tensor] = None):
        """Replay the step.  `y0_host` (optional) is copied into the pinned input buffer first; otherwise whatever the
        caller wrote into `self.y0_host` is used.  Asynchronous: call `sync()` before reading `self.host[...]`."""
        if y0_host is not None:
            self.y0_host.copy_(y0_host)
        self.graph.replay()

    def sync(self):
        """Wait for the replay and return the pinned host buffers.  Raises torchdiffeq's solver asserts if the solve failed
        on the device (status mailbox; a plain host-memory read after the synchronisation)."""
        torch.cuda.current_stream(self.device).synchronize()
        _api.check_status()
        return self.host


class GraphedSolvePipeline:
    """`depth` GraphedSolveSteps in flight: while slot i computes, slot i+1's input is already crossing PCIe.

        pipe = GraphedSolvePipeline(ode_fn, B, t, depth=2, method='dopri5', rtol=1e-5, atol=1e-5, adjoint=False)
        for s in pipe.slots: s.grad_traj.copy_(upstream)
        pipe.submit(y0_host)                 # H2D on the slot's stream -> wait for the previous slot's kernels -> graph
        ...                                  #   (fwd + bwd, one cudaGraphLaunch) -> D2H on the slot's stream
        out = pipe.result()                  # oldest step in flight: wait for ITS D2H only, status check, pinned buffers

    Every step still does its own H2D and D2H; what changes against run() + sync() per step is that the host does not wait
    for step i before it issues step i+1.  The solver kernels of consecutive slots are ordered by an event (they are
    cooperative launches that want the whole GPU, and with the fused multi-GPU gradient exchange every rank must run them
    in the same order); only the copies overlap them.  Results come back in submission order."""

    def __init__(self, func, batch: int, t: torch.Tensor, *, depth: int = 2, **kw):
        assert depth >= 1
        kw.pop("copies_in_graph", None)
        self.slots = [GraphedSolveStep(func, batch, t, copies_in_graph=False, **kw) for _ in range(depth)]
        dev = self.slots[0].device
        self.device = dev
        self._streams = [torch.cuda.Stream(device=dev) for _ in range(depth)]
        self._computed = [torch.cuda.Event() for _ in range(depth)]   # the slot's kernels are done
        self._done = [torch.cuda.Event() for _ in range(depth)]       # ... and its D2H copies
        self._head, self._inflight, self._last = 0, [], None
        for st in self._streams:
            st.wait_stream(torch.cuda.current_stream(dev))

    def next_input(self) -> torch.Tensor:
        """The pinned input buffer the next submit() will send: write the batch there to save one host copy."""
        return self.slots[self._head].y0_host

    def submit(self, y0_host: Optional[torch.Tensor] = None) -> int:
        if len(self._inflight) == len(self.slots):
            raise RuntimeError("pipeline full: call result() before submitting another step")
        i = self._head
        slot, st = self.slots[i], self._streams[i]
        if y0_host is not None:
            slot.y0_host.copy_(y0_host)
        with torch.cuda.stream(st):
            slot.copy_in()
            if self._last is not None and self._last != i:
                st.wait_event(self._computed[self._last])
            slot.graph.replay()
            self._computed[i].record(st)
            slot.copy_out()
            self._done[i].record(st)
        self._last = i
        self._inflight.append(i)
        self._head = (i + 1) % len(self.slots)
        return i

    def result(self):
        """Block until the oldest submitted step has landed in its pinned buffers and return them."""
        i = self._inflight.pop(0)
        self._done[i].synchronize()
        _api.check_status()
        return self.slots[i].host

    def drain(self):
        out = None
        while self._inflight:
            out = self.result()
        return out
