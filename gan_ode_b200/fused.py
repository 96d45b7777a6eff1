"""The latent-motion sampler fused around the solve — OPT-IN.

What the reference does per `sample_z_video` (models/mocogan.py:259-269 + models/mocogan_ode.py:133-148):

    z_content = numpy normal -> repeat over frames -> H2D            (n*T, dc)
    x = torch.randn(n, D).cuda(); x = self.linear(x)                 randn + 4 launches (2 addmm, 2 leaky_relu)
    z_m = odeint(ode_fn, x, linspace(0,1,T), method='rk4')           (T, n, D)
    z_m = z_m.transpose(0, 1).reshape(-1, D)                         a copy
    z = torch.cat([z_content, z_motion], dim=1)                      another copy

and per `sample_images(n)` (models/mocogan.py:287-295) all of that for n*T*2 videos, of whose n*T*2*T rows n are kept.

Here: ONE kernel launch (`gode_rk4_sampler_fwd`) draws x (Philox, keyed by global trajectory index), applies the pre-MLP,
integrates, and writes the frame-major codes straight into the motion columns of `z`; the content columns are filled by one
broadcast copy.  `sample_images` solves only the trajectories whose rows are kept — bit-identical to solving all of them and
indexing, because the noise is a function of (seed, trajectory index).  The backward is the same continuous adjoint as
`odeint_adjoint(..., method='rk4')` (`gode_rk4_adjoint_bwd_strided`, reading codes and upstream gradient in place inside z and
grad_z), followed by the pre-MLP's backward on the stored noise (a few small PyTorch ops).

Opt-in because the random numbers differ from the reference's: x comes from Philox instead of torch's generator (the numpy
draws for z_content and for the kept rows are made exactly as the reference makes them).  Use `FusedLatentSampler(gen)` to get
drop-in `sample_z_video / sample_videos / sample_images` for a reference generator (or `.install()` to patch them in).
"""
from __future__ import annotations

import importlib

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib

_api = importlib.import_module(__package__ + ".odeint")

__all__ = ["fused_sample_z", "FusedLatentSampler", "recognise_pre_mlp"]


def recognise_pre_mlp(linear):
    """(Wa, ba, Wb, bb, slope) of the reference's pre-MLP (models/mocogan_ode.py:123-131:
    Sequential(Linear(D,64), LeakyReLU(0.2), Linear(64,D), LeakyReLU(0.2))), or None for nn.Identity (:36-37)."""
    if isinstance(linear, nn.Identity):
        return None
    ok = (isinstance(linear, nn.Sequential) and len(linear) == 4 and isinstance(linear[0], nn.Linear)
          and isinstance(linear[1], nn.LeakyReLU) and isinstance(linear[2], nn.Linear) and isinstance(linear[3], nn.LeakyReLU)
          and linear[0].bias is not None and linear[2].bias is not None and linear[0].out_features == 64
          and linear[2].in_features == 64 and linear[0].in_features == linear[2].out_features
          and linear[1].negative_slope == linear[3].negative_slope)
    if not ok:
        raise NotImplementedError("the fused sampler knows the reference's pre-MLP (models/mocogan_ode.py:123-131) or "
                                  "nn.Identity; got {!r}".format(linear))
    return linear[0].weight, linear[0].bias, linear[2].weight, linear[2].bias, float(linear[1].negative_slope)


class _FusedSampler(torch.autograd.Function):
    @staticmethod
    def forward(ctx, meta, content, W1, b1, W2, b2, *pre):
        L = _lib.lib()
        B, T, D, H = meta["B"], meta["T"], W1.shape[1], W1.shape[0]
        dev = W1.device
        ws = [_api._f32c(w) for w in (W1, b1, W2, b2)]
        pw = [_api._f32c(w) for w in pre]
        dc = 0 if content is None else content.shape[1]
        C_ = dc + D
        z = torch.empty((B * T, C_), dtype=torch.float32, device=dev)
        if dc:   # models/mocogan.py:249-257: the content code of a video, repeated over its frames
            z.view(B, T, C_)[:, :, :dc].copy_(content.view(B, 1, dc).expand(B, T, dc))
        noise = torch.empty((B, D), dtype=torch.float32, device=dev)
        dt = meta["dt"]
        ids = meta["traj_ids"]
        rc = L.gode_rk4_sampler_fwd(*( [w.data_ptr() for w in pw] if pw else [None] * 4), meta["slope"], 64 if pw else 0,
                                    *[w.data_ptr() for w in ws], dt.ctypes.data, 0, B, D, H, T, meta["seed"], meta["traj_offset"],
                                    None if ids is None else ids.data_ptr(), _lib.LAYOUT_BTD, z.data_ptr() + 4 * dc, C_,
                                    noise.data_ptr(), _api._stream())
        _lib.check(rc, "gode_rk4_sampler_fwd")
        ctx.meta, ctx.dc = meta, dc
        ctx.save_for_backward(z, noise, *ws, *pw)
        return z

    @staticmethod
    @_api._bwd_on_device
    def backward(ctx, grad_z):
        L = _lib.lib()
        z, noise, W1, b1, W2, b2, *pw = ctx.saved_tensors
        meta, dc = ctx.meta, ctx.dc
        B, T, D, H = meta["B"], meta["T"], W1.shape[1], W1.shape[0]
        C_ = dc + D
        g = grad_z if (grad_z.dtype is torch.float32 and grad_z.is_contiguous() and not (grad_z.data_ptr() & 15)) else \
            grad_z.detach().float().contiguous().clone()
        n_param = L.gode_param_count(D, H)
        grad_p = torch.empty(n_param, dtype=torch.float32, device=z.device)
        grad_y0 = torch.empty((B, D), dtype=torch.float32, device=z.device)
        ws_bytes = L.gode_bwd_workspace_bytes(B, D, H)
        wsp = _api._workspace(z.device, ws_bytes)
        dt = meta["dt"]
        rc = L.gode_rk4_adjoint_bwd_strided(z.data_ptr() + 4 * dc, C_, g.data_ptr() + 4 * dc, C_, W1.data_ptr(), b1.data_ptr(),
                                            W2.data_ptr(), b2.data_ptr(), dt.ctypes.data, 0, B, D, H, T, grad_y0.data_ptr(),
                                            grad_p.data_ptr(), wsp.data_ptr(), ws_bytes, _api._stream())
        _lib.check(rc, "gode_rk4_adjoint_bwd_strided")
        needs = ctx.needs_input_grad
        gW1, gb1, gW2, gb2 = _api._split_params(grad_p, D, H, needs[2:6])
        gpre = [None] * len(pw)
        if pw and any(needs[6:10]):   # backward of y0 = linear(x) on the stored noise: four small PyTorch ops each way
            with torch.enable_grad():
                ps = [w.detach().requires_grad_(True) for w in pw]
                y0 = F.leaky_relu(F.linear(F.leaky_relu(F.linear(noise, ps[0], ps[1]), meta["slope"]), ps[2], ps[3]), meta["slope"])
                gs = torch.autograd.grad(y0, ps, grad_y0)
            gpre = [gk if need else None for gk, need in zip(gs, needs[6:10])]
        return (None, None, gW1, gb1, gW2, gb2, *gpre)


def fused_sample_z(linear, ode_fn, num_samples, video_len, *, content=None, seed=None, traj_offset=0, traj_ids=None):
    """Frame-major latent codes for `num_samples` videos in one launch.  Returns z of shape (num_samples * video_len, dc + D):
    row b*T + j = [content[b] | motion code of trajectory b at frame j] (dc = 0 without `content`), i.e. what
    `torch.cat([z_content, sample_z_m(...)], dim=1)` builds in the reference.  `traj_ids` (int64 tensor, num_samples): global
    trajectory indices of the rows when only a subset of a larger batch is solved."""
    W1, b1, W2, b2 = _api.recognise_field(ode_fn)
    pre = recognise_pre_mlp(linear)
    D, H = W1.shape[1], W1.shape[0]
    if (D, H) != (16, 16):
        raise NotImplementedError("the fused sampler exists for the reference shape D = H = 16")
    weights = (W1, b1, W2, b2) + (tuple(pre[:4]) if pre else ())
    anchor = W1
    _api._require_cuda(anchor, what="ode_fn", weights=weights)
    if content is not None:
        if content.dim() != 2 or content.shape[0] != num_samples or content.shape[1] % 2:
            raise ValueError("content must be (num_samples, dc) with even dc")
        _api._require_cuda(anchor, content, what="ode_fn / content")
        content = _api._f32c(content)
    if traj_ids is not None:
        traj_ids = traj_ids.to(device=anchor.device, dtype=torch.int64).contiguous()
        if traj_ids.numel() != num_samples:
            raise ValueError("traj_ids must have num_samples entries")
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())     # reproducible under torch.manual_seed
    t = torch.linspace(0, 1, video_len).float()                 # models/mocogan_ode.py:143
    dt = _api._rk4_dt(t, {}, anchor.device)
    meta = dict(B=int(num_samples), T=int(video_len), dt=dt, seed=int(seed) & 0xFFFFFFFFFFFFFFFF, traj_offset=int(traj_offset),
                traj_ids=traj_ids, slope=pre[4] if pre else 0.0)
    with _api._on_device(anchor.device):
        return _FusedSampler.apply(meta, content, W1, b1, W2, b2, *(pre[:4] if pre else ()))


class FusedLatentSampler:
    """Drop-in `sample_z_video / sample_videos / sample_images` for a reference generator (models/mocogan_ode.py::
    VideoGenerator* on models/mocogan.py::VideoGenerator; dim_z_category = 0 as in every reference script).  `install()` patches
    the generator's methods; `uninstall()` restores them."""

    def __init__(self, gen):
        if getattr(gen, "dim_z_category", 0):
            raise NotImplementedError("categorical codes are not used by the reference's ODE scripts (dim_z_category = 0)")
        self.gen = gen
        self._saved = None

    def _content(self, n):
        # models/mocogan.py:249-257, the same numpy draw; the repeat over frames happens inside the fused buffer fill
        c = np.random.normal(0, 1, (n, self.gen.dim_z_content)).astype(np.float32)
        return torch.from_numpy(c).to(next(self.gen.ode_fn.parameters()).device, non_blocking=True)

    def sample_z_video(self, num_samples, video_len=None):
        T = video_len if video_len is not None else self.gen.video_length
        z = fused_sample_z(self.gen.linear, self.gen.ode_fn, num_samples, T, content=self._content(num_samples))
        return z, np.zeros(num_samples)

    def sample_videos(self, num_samples, video_len=None):      # models/mocogan.py:271-285
        T = video_len if video_len is not None else self.gen.video_length
        z, labels = self.sample_z_video(num_samples, T)
        h = self.gen.main(z.view(z.size(0), z.size(1), 1, 1))
        h = h.view(h.size(0) // T, T, self.gen.n_channels, h.size(3), h.size(3))
        return h.permute(0, 2, 1, 3, 4), torch.from_numpy(labels).to(z.device)

    def sample_image_codes(self, num_samples, seed=None):
        """The z rows `sample_images` feeds to the frame generator: rows j of sample_z_video(num_samples * T * 2), computed
        by solving only the trajectories those rows belong to."""
        T = self.gen.video_length
        n_all = num_samples * T * 2
        content = self._content(n_all)
        j = np.sort(np.random.choice(n_all * T, num_samples, replace=False)).astype(np.int64)   # models/mocogan.py:290
        b, f = torch.from_numpy(j // T).to(content.device), torch.from_numpy(j % T).to(content.device)
        codes = fused_sample_z(self.gen.linear, self.gen.ode_fn, num_samples, T, seed=seed, traj_ids=b)   # (n*T, D)
        D = codes.shape[1]
        picked = codes.view(num_samples, T, D)[torch.arange(num_samples, device=codes.device), f]
        return torch.cat([content[b], picked], dim=1), j

    def sample_images(self, num_samples):                      # models/mocogan.py:287-295
        z, _ = self.sample_image_codes(num_samples)
        return self.gen.main(z.view(z.size(0), z.size(1), 1, 1)), None

    def install(self):
        g = self.gen
        self._saved = {k: g.__dict__.get(k) for k in ("sample_z_video", "sample_videos", "sample_images")}
        g.sample_z_video, g.sample_videos, g.sample_images = self.sample_z_video, self.sample_videos, self.sample_images
        return self

    def uninstall(self):
        for k, v in (self._saved or {}).items():
            if v is None:
                self.gen.__dict__.pop(k, None)
            else:
                setattr(self.gen, k, v)
        self._saved = None
