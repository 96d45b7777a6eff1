"""Host-side mirrors of the reference's vector-field modules — the structures `recognise_field` / `recognise_sde` accept.

These are plain `nn.Module`s with the reference's constructor signature, attribute names and `state_dict` keys, so a
checkpoint written by the reference scripts (mnist_moco_ode.py:175-182) loads into them unchanged.  The product path
never calls their `forward`: the solver kernels read the four (eight) weight tensors directly.  `bench.py` and the data-
parallel harness build their synthetic fields from here (no import of `oracle/` on the product arm).
"""
from __future__ import annotations

import torch
import torch.nn as nn

__all__ = ["ODEFunc", "SDEFunc", "make_field"]


class ODEFunc(nn.Module):
    """models/mocogan_ode.py:6-17 (dup models/mocogan_ode_rnn.py:6-17): f(t, x) = W2 tanh(W1 x + b1) + b2, t ignored;
    parameters() order fn.0.weight, fn.0.bias, fn.2.weight, fn.2.bias."""

    def __init__(self, dim, dim_hidden):
        super().__init__()
        self.fn = nn.Sequential(nn.Linear(dim, dim_hidden), nn.Tanh(), nn.Linear(dim_hidden, dim))

    def forward(self, t, x):
        return self.fn(x)


class SDEFunc(nn.Module):
    """models/mocogan_sde.py:6-27: drift and diffusion are two independent ODEFunc-shaped MLPs; diagonal Ito noise."""

    noise_type = "diagonal"
    sde_type = "ito"

    def __init__(self, dim, dim_hidden):
        super().__init__()
        self.drift_fn = nn.Sequential(nn.Linear(dim, dim_hidden), nn.Tanh(), nn.Linear(dim_hidden, dim))
        self.diffusion_fn = nn.Sequential(nn.Linear(dim, dim_hidden), nn.Tanh(), nn.Linear(dim_hidden, dim))

    def f(self, t, x):
        return self.drift_fn(x)

    def g(self, t, x):
        return self.diffusion_fn(x)


def make_field(D=16, H=16, seed=0, scale=1.0, device="cpu"):
    """ODEFunc(D, H) with PyTorch's default nn.Linear init under torch.manual_seed(seed) (BASELINE.md §4's synthetic
    weights); `scale` multiplies every parameter (the stiffer dopri5 variant)."""
    torch.manual_seed(seed)
    f = ODEFunc(D, H)
    if scale != 1.0:
        with torch.no_grad():
            for p in f.parameters():
                p.mul_(scale)
    return f.to(device)
