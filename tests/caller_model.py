"""Stand-in for the reference's caller of the hot path, for tests that must run where /root/reference is absent
(the GPU box).  Restates ONLY the latent-motion part of VideoGeneratorMNISTODE (models/mocogan_ode.py:114-148):
pre-MLP `linear` (:123-131), `ode_fn` (:118-121) and `sample_z_m` (:133-148), with the solver imported exactly the
way the reference imports it (`from torchdiffeq import odeint_adjoint as odeint`, models/mocogan_ode.py:4) so the
shim is what gets exercised.  tests/test_reference_dropin_cpu.py checks this stand-in against the real reference
file (same state_dict keys, same output for the same seed) whenever /root/reference exists."""
import torch
import torch.nn as nn


class ODEFunc(nn.Module):  # models/mocogan_ode.py:6-17
    def __init__(self, dim, dim_hidden):
        super().__init__()
        self.fn = nn.Sequential(nn.Linear(dim, dim_hidden), nn.Tanh(), nn.Linear(dim_hidden, dim))

    def forward(self, t, x):
        return self.fn(x)


class LatentMotionODE(nn.Module):
    def __init__(self, dim_z_motion=16, video_length=16, dim_hidden=None):
        super().__init__()
        self.dim_z_motion, self.video_length = dim_z_motion, video_length
        self.ode_fn = ODEFunc(dim=dim_z_motion, dim_hidden=dim_hidden or dim_z_motion)
        self.linear = nn.Sequential(nn.Linear(dim_z_motion, 64), nn.LeakyReLU(0.2),
                                    nn.Linear(64, dim_z_motion), nn.LeakyReLU(0.2))

    def sample_z_m(self, num_samples, video_len=None, noise=None):
        from torchdiffeq import odeint_adjoint as odeint  # resolved through sys.modules (shim or oracle)
        video_len = video_len if video_len is not None else self.video_length
        x = torch.randn(num_samples, self.dim_z_motion) if noise is None else noise
        x = x.to(next(self.parameters()).device)
        x = self.linear(x)
        z_m_t = odeint(self.ode_fn, x, torch.linspace(0, 1, video_len).float(), method='rk4')
        return z_m_t.transpose(0, 1).reshape(-1, self.dim_z_motion)


class LatentMotionODERNN(nn.Module):
    """ODE-RNN sampler: models/mocogan_ode_rnn.py:21-54 (+ models/mocogan.py:198 GRUCell, :297-301 noise sources).
    Per frame: h' = odeint(ode_fn, h, tensor([0,1]))[-1] with torchdiffeq DEFAULTS (dopri5, rtol 1e-7, atol 1e-9),
    then h = GRUCell(e_t, h').  `h0` / `eps` may be injected so runs are repeatable."""

    def __init__(self, dim_z_motion=16, video_length=16):
        super().__init__()
        self.dim_z_motion, self.video_length = dim_z_motion, video_length
        self.recurrent = nn.GRUCell(dim_z_motion, dim_z_motion)
        self.ode_fn = ODEFunc(dim=dim_z_motion, dim_hidden=dim_z_motion)

    def sample_z_m(self, num_samples, video_len=None, h0=None, eps=None, **solver_kw):
        from torchdiffeq import odeint_adjoint as odeint
        video_len = video_len if video_len is not None else self.video_length
        dev = next(self.parameters()).device
        h_t = [torch.randn(num_samples, self.dim_z_motion, device=dev) if h0 is None else h0.to(dev)]
        for frame_num in range(video_len):
            e_t = torch.randn(num_samples, self.dim_z_motion, device=dev) if eps is None else eps[frame_num].to(dev)
            h_t_prime = odeint(self.ode_fn, h_t[-1], torch.tensor([0, 1]).float(), **solver_kw)[-1]
            h_t.append(self.recurrent(e_t, h_t_prime))
        z_m_t = [h_k.view(-1, 1, self.dim_z_motion) for h_k in h_t]
        return torch.cat(z_m_t[1:], dim=1).view(-1, self.dim_z_motion)
