"""CPU restatement of the reference's latent-motion modules and sampler loops.

TEST INFRASTRUCTURE ONLY (see oracle/torchdiffeq_restatement.py header; same rules).
Each piece cites the reference lines it follows; the solver underneath is the restatement in
this directory, because the reference's own solver dependency is absent here.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import torchdiffeq_restatement as tdq
from . import torchsde_restatement as tsde


class ODEFunc(nn.Module):
    """models/mocogan_ode.py:6-17 (dup models/mocogan_ode_rnn.py:6-17): autonomous
    f(t, x) = W2 tanh(W1 x + b1) + b2 with parameter order fn.0.weight, fn.0.bias, fn.2.weight, fn.2.bias."""

    def __init__(self, dim, dim_hidden):
        super().__init__()
        self.fn = nn.Sequential(nn.Linear(dim, dim_hidden), nn.Tanh(), nn.Linear(dim_hidden, dim))

    def forward(self, t, x):
        return self.fn(x)


class SDEFunc(nn.Module):
    """models/mocogan_sde.py:6-27: two independent ODEFunc-shaped MLPs, diagonal Ito noise."""

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


def frame_major(traj: torch.Tensor) -> torch.Tensor:
    """models/mocogan_ode.py:146: (T, B, D) -> (B*T, D), row b*T + j = trajectory b at frame j."""
    return traj.transpose(0, 1).reshape(-1, traj.shape[-1])


def sample_z_m_ode(ode_fn, x, video_len, method="rk4", adjoint=True, **kw):
    """models/mocogan_ode.py:142-146 after the pre-MLP: solve on linspace(0,1,T), emit frame-major codes."""
    t = torch.linspace(0, 1, video_len).float()
    solve = tdq.odeint_adjoint if adjoint else tdq.odeint
    return frame_major(solve(ode_fn, x, t, method=method, **kw))


def sample_z_m_odernn(ode_fn, gru: nn.GRUCell, h0, eps, adjoint=True, **kw):
    """models/mocogan_ode_rnn.py:40-54: per frame h' = odeint(ode_fn, h, [0,1])[-1] (torchdiffeq defaults:
    dopri5, rtol 1e-7, atol 1e-9 unless overridden in **kw), h = GRUCell(e_t, h').  `eps` is (T, B, D): the
    per-frame noise the reference draws at :46 (models/mocogan.py:300-301), supplied so runs are repeatable.
    Output (B*T, D) frame-major within sample (:51-52)."""
    solve = tdq.odeint_adjoint if adjoint else tdq.odeint
    h = h0
    hs = []
    for e_t in eps:
        h_prime = solve(ode_fn, h, torch.tensor([0, 1]).float(), **kw)[-1]
        h = gru(e_t, h_prime)
        hs.append(h)
    D = h0.shape[-1]
    return torch.cat([h_k.view(-1, 1, D) for h_k in hs], dim=1).view(-1, D)


def sample_z_m_sde(sde, x, video_len, increments, dt=2.5e-2):
    """models/mocogan_sde.py:57-61 given the Brownian increments of the 41-step grid."""
    ts = torch.linspace(0, 1, video_len).float()
    bm = tsde.TableBrownian(increments)
    return frame_major(tsde.sdeint(sde, x, ts, bm=bm, method="euler", dt=dt))
