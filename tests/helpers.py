"""Shared helpers for the parity tests (seeded synthetic inputs as BASELINE.md §4 prescribes)."""
import copy

import torch

from oracle.latent_motion import ODEFunc, SDEFunc  # noqa: F401  (the oracle's restatement of the reference modules)


def make_field(D=16, H=16, seed=0, scale=1.0, device="cpu", dtype=torch.float32):
    """Freshly constructed ODEFunc(D, H) with PyTorch default nn.Linear init under torch.manual_seed(seed);
    `scale` multiplies every parameter (the 'stiffer' dopri5 variant of SURVEY §8d)."""
    torch.manual_seed(seed)
    f = ODEFunc(D, H)
    if scale != 1.0:
        with torch.no_grad():
            for p in f.parameters():
                p.mul_(scale)
    return f.to(device=device, dtype=dtype)


def clone_to(f, device, dtype=torch.float32):
    g = copy.deepcopy(f).to(device=device, dtype=dtype)
    for p in g.parameters():
        p.grad = None
    return g


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b|  (norm-wise relative error, b is the reference)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def grads_of(loss, tensors):
    return torch.autograd.grad(loss, tensors)


def elem_close(a: torch.Tensor, ref: torch.Tensor, atol_rel=1e-5, rtol=1e-4):
    """Element-wise bar beside the norm-wise one: |a - ref| <= atol_rel * max|ref| + rtol * |ref| for EVERY entry.
    Returns (ok, worst) where worst = max over entries of |a-ref| / (atol + rtol |ref|) (<= 1 passes)."""
    a = a.detach().double().cpu()
    ref = ref.detach().double().cpu()
    bound = atol_rel * ref.abs().max().clamp_min(1e-30) + rtol * ref.abs()
    worst = float(((a - ref).abs() / bound).max())
    return worst <= 1.0, worst


def rowwise_rel_err(a: torch.Tensor, ref: torch.Tensor) -> float:
    """Per-trajectory relative error: max over rows (all leading dims) of max|a-ref|_row / max|ref|_row.  A trajectory
    whose values are small against the batch maximum is still held to its own scale."""
    a = a.detach().double().cpu().reshape(-1, a.shape[-1])
    ref = ref.detach().double().cpu().reshape(-1, ref.shape[-1])
    num = (a - ref).abs().amax(dim=1)
    den = ref.abs().amax(dim=1).clamp_min(1e-30)
    return float((num / den).max())
