"""configs[3] end to end: a data-parallel MoCoGAN-ODE training step on synthetic UCF-shaped clips.

    python scripts/train_dp_harness.py                                     # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29521 \
        scripts/train_dp_harness.py [--iters K] [--batch B] [--fp32]

The loop is the reference's train() (ucf_moco_ode.py:113-163): per iteration two discriminator rounds (image + video
discriminator on real clips and on generator samples drawn under no_grad) and one generator step whose backward runs through
the latent-motion ODE; three Adam optimisers (lr 2e-4, betas (0.5, 0.999), weight decay 1e-5, ucf_moco_ode.py:86-88);
BCE-with-logits.  The nets are the reference's OWN classes whenever /root/reference is mounted —
`models.mocogan_ode.VideoGenerator(3, 50, 0, 64, 16, dim_hidden=256)` (its sample_z_m / sample_z_video / sample_images /
sample_videos run unmodified; `dim_hidden` is passed because ucf_moco_ode.py:80 forgets it),
`models.mocogan.VideoDiscriminator`, `models.mocogan.PatchImageDiscriminator` — and otherwise (the GPU box has no
/root/reference) stand-ins with the same layers, state_dict keys and shapes (tests/test_harness.py checks them against the
reference classes).  The data are random clips (B, 3, 16, 64, 64).  `from torchdiffeq import odeint_adjoint` inside the
generator resolves to gan_ode_b200 through the shim: D=64 / H=256 motion ODE, tcgen05 forward + tcgen05 adjoint in bf16 mode;
the ODE parameter gradient is all-reduced by this library, the conv nets by DDP.  Reports iterations/s and the share of the
step spent in the ODE solves (CUDA events).
"""
import argparse
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference"
DZM, DH, DZC, T = 64, 256, 50, 16


# ---- stand-ins for the GPU box: same layers / keys / shapes as models/mocogan.py:185-215, :66-93, :127-163 ----------------
class _ODEFunc(nn.Module):  # models/mocogan_ode.py:6-17
    def __init__(self, dim, dim_hidden):
        super().__init__()
        self.fn = nn.Sequential(nn.Linear(dim, dim_hidden), nn.Tanh(), nn.Linear(dim_hidden, dim))

    def forward(self, t, x):
        return self.fn(x)


def _conv_stack(conv, bn, chans, first, last, n_noise):
    """[Noise, conv, (bn), LeakyReLU] blocks as the reference lays them out (the parameter-free Noise modules only shift the
    Sequential indices; kept as Identity so that state_dict keys match)."""
    layers = []
    for i in range(len(chans) - 1):
        if i < n_noise:
            layers.append(nn.Identity())
        layers.append(conv(chans[i], chans[i + 1], *first))
        if i > 0:
            layers.append(bn(chans[i + 1]))
        layers.append(nn.LeakyReLU(0.2, inplace=True))
    layers.append(last)
    return nn.Sequential(*layers)


class StandInGenerator(nn.Module):
    """models/mocogan_ode.py:20-54 on models/mocogan.py:185-295: members recurrent / main / ode_fn / linear, same samplers."""

    def __init__(self, n_channels=3, dim_z_content=DZC, dim_z_motion=DZM, video_length=T, dim_hidden=DH, ngf=64):
        super().__init__()
        self.n_channels, self.dim_z_content, self.dim_z_motion, self.video_length = n_channels, dim_z_content, dim_z_motion, video_length
        self.recurrent = nn.GRUCell(dim_z_motion, dim_z_motion)   # constructed by the base class, unused by the ODE sampler
        c = [dim_z_content + dim_z_motion, ngf * 8, ngf * 4, ngf * 2, ngf]
        layers = []
        for i in range(4):
            layers += [nn.ConvTranspose2d(c[i], c[i + 1], 4, 1 if i == 0 else 2, 0 if i == 0 else 1, bias=False),
                       nn.BatchNorm2d(c[i + 1]), nn.ReLU(True)]
        layers += [nn.ConvTranspose2d(ngf, n_channels, 4, 2, 1, bias=False), nn.Tanh()]
        self.main = nn.Sequential(*layers)
        self.ode_fn = _ODEFunc(dim_z_motion, dim_hidden)
        self.linear = nn.Sequential(nn.Linear(dim_z_motion, 64), nn.LeakyReLU(0.2), nn.Linear(64, dim_z_motion), nn.LeakyReLU(0.2))

    def sample_z_m(self, num_samples, video_len=None):     # models/mocogan_ode.py:39-54
        from torchdiffeq import odeint_adjoint as odeint
        video_len = video_len if video_len is not None else self.video_length
        dev = self.linear[0].weight.device
        x = self.linear(torch.randn(num_samples, self.dim_z_motion).to(dev))
        z = odeint(self.ode_fn, x, torch.linspace(0, 1, video_len).float(), method='rk4')
        return z.transpose(0, 1).reshape(-1, self.dim_z_motion)

    def sample_z_video(self, num_samples, video_len=None):  # models/mocogan.py:249-269 (dim_z_category = 0)
        video_len = video_len if video_len is not None else self.video_length
        dev = self.linear[0].weight.device
        content = np.repeat(np.random.normal(0, 1, (num_samples, self.dim_z_content)).astype(np.float32), video_len, axis=0)
        return torch.cat([torch.from_numpy(content).to(dev), self.sample_z_m(num_samples, video_len)], dim=1), None

    def sample_videos(self, num_samples, video_len=None):   # models/mocogan.py:271-285
        video_len = video_len if video_len is not None else self.video_length
        z, _ = self.sample_z_video(num_samples, video_len)
        h = self.main(z.view(z.size(0), z.size(1), 1, 1))
        h = h.view(h.size(0) // video_len, video_len, self.n_channels, h.size(3), h.size(3))
        return h.permute(0, 2, 1, 3, 4), None

    def sample_images(self, num_samples):                   # models/mocogan.py:287-295
        z, _ = self.sample_z_video(num_samples * self.video_length * 2)
        j = np.sort(np.random.choice(z.size(0), num_samples, replace=False)).astype(np.int64)
        z = z[j, ::]
        return self.main(z.view(z.size(0), z.size(1), 1, 1)), None


class StandInPatchImageD(nn.Module):   # models/mocogan.py:66-93
    def __init__(self, n_channels=3, ndf=64):
        super().__init__()
        self.main = _conv_stack(lambda a, b, *k: nn.Conv2d(a, b, *k, bias=False), nn.BatchNorm2d, [n_channels, ndf, ndf * 2, ndf * 4],
                                (4, 2, 1), nn.Conv2d(ndf * 4, 1, 4, 2, 1, bias=False), 3)
        self.main.insert(len(self.main) - 1, nn.Identity())   # the Noise in front of the last conv

    def forward(self, x):
        return self.main(x).squeeze(), None


class StandInVideoD(nn.Module):        # models/mocogan.py:127-163
    def __init__(self, n_channels=3, ndf=64):
        super().__init__()
        k = dict(stride=(1, 2, 2), padding=(0, 1, 1), bias=False)
        self.main = _conv_stack(lambda a, b, *_: nn.Conv3d(a, b, 4, **k), nn.BatchNorm3d, [n_channels, ndf, ndf * 2, ndf * 4, ndf * 8],
                                (), nn.Conv3d(ndf * 8, 1, 4, 1, 0, bias=False), 4)

    def forward(self, x):
        return self.main(x).squeeze(), None


def reference_modules():
    """The reference's own model modules (models.mocogan, models.mocogan_ode), importable once a `torchdiffeq` is in
    sys.modules; None when /root/reference is not mounted."""
    if not os.path.isdir(os.path.join(REF, "models")):
        return None
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import importlib
    return importlib.import_module("models.mocogan"), importlib.import_module("models.mocogan_ode")


def build_nets(use_reference=None):
    ref = reference_modules() if use_reference in (None, True) else None
    if use_reference and ref is None:
        raise RuntimeError("/root/reference is not mounted")
    if ref is not None:
        mocogan, mocogan_ode = ref
        gen = mocogan_ode.VideoGenerator(3, DZC, 0, DZM, T, dim_hidden=DH)       # ucf_moco_ode.py:80 (+ dim_hidden, Appendix C)
        return gen, mocogan.PatchImageDiscriminator(3), mocogan.VideoDiscriminator(3), "reference classes (models/mocogan*.py)"
    return StandInGenerator(), StandInPatchImageD(), StandInVideoD(), "stand-ins (no /root/reference on this box)"


def run_harness(iters=10, batch=32, precision="bf16", use_reference=None, device=None, solver_module=None, warmup=3):
    """One process of the data-parallel harness.  `solver_module`: what `import torchdiffeq` resolves to (default: the
    gan_ode_b200 shim; the CPU test passes the oracle).  Returns the result dict (rank 0) or None."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    on_gpu = device is None or torch.device(device).type == "cuda"
    if on_gpu:
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
    else:
        dev = torch.device("cpu")
    prev_mod = sys.modules.get("torchdiffeq")
    if solver_module is None:
        import gan_ode_b200 as gode
        gode.install_shims()
        gode.config.layout, gode.config.precision = "btd", precision
        if world > 1:
            gode.config.grad_allreduce = True
    else:
        sys.modules["torchdiffeq"] = solver_module
    try:
        torch.manual_seed(0)
        np.random.seed(0)
        gen, dimg, dvid, which = build_nets(use_reference)
        gen, dimg, dvid = gen.to(dev), dimg.to(dev), dvid.to(dev)
        ode_params = list(gen.ode_fn.parameters())
        g_train = gen
        if world > 1:  # conv nets: DDP; ODE parameters: this library's all-reduce (they must not be in a DDP bucket as well)
            dimg = nn.parallel.DistributedDataParallel(dimg, device_ids=[local], broadcast_buffers=False)
            dvid = nn.parallel.DistributedDataParallel(dvid, device_ids=[local], broadcast_buffers=False)
            for p in gen.parameters():
                dist.broadcast(p.data, 0)
            ode_ids = {id(p) for p in ode_params}
            hooks = [p for p in gen.parameters() if id(p) not in ode_ids]

            def allreduce_rest():   # the generator's non-ODE gradients (its samplers are methods, so DDP cannot wrap it)
                flat = torch.cat([p.grad.reshape(-1) for p in hooks if p.grad is not None])
                dist.all_reduce(flat)
                flat /= world
                o = 0
                for p in hooks:
                    if p.grad is not None:
                        p.grad.copy_(flat[o:o + p.numel()].view_as(p))
                        o += p.numel()
        adam = dict(lr=2e-4, betas=(0.5, 0.999), weight_decay=1e-5)
        opt_g = torch.optim.Adam(gen.parameters(), **adam)
        opt_i, opt_v = torch.optim.Adam(dimg.parameters(), **adam), torch.optim.Adam(dvid.parameters(), **adam)
        bce = nn.BCEWithLogitsLoss()
        B = batch
        ode_ev = []
        orig_sample_z_m = gen.sample_z_m

        def timed_sample_z_m(*a, **k):   # CUDA events around the reference's own sample_z_m
            if not on_gpu:
                return orig_sample_z_m(*a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            z = orig_sample_z_m(*a, **k)
            e1.record()
            ode_ev.append((e0, e1))
            return z
        gen.sample_z_m = timed_sample_z_m
        losses = {}

        def iteration():                  # ucf_moco_ode.py:113-163
            for _ in range(2):
                real = torch.randn(B, 3, T, 64, 64, device=dev)
                opt_i.zero_grad()
                pr, _ = dimg(real[:, :, 0])
                with torch.no_grad():
                    fake, _ = g_train.sample_images(B)
                pf, _ = dimg(fake)
                li = bce(pr, torch.ones_like(pr)) + bce(pf, torch.zeros_like(pf))
                li.backward()
                opt_i.step()
                opt_v.zero_grad()
                pr, _ = dvid(real)
                with torch.no_grad():
                    fake, _ = g_train.sample_videos(B)
                pf, _ = dvid(fake)
                lv = bce(pr, torch.ones_like(pr)) + bce(pf, torch.zeros_like(pf))
                lv.backward()
                opt_v.step()
            opt_g.zero_grad()
            fv, _ = g_train.sample_videos(B)
            fi, _ = g_train.sample_images(B)
            pv, _ = dvid(fv)
            pi, _ = dimg(fi)
            lg = bce(pv, torch.ones_like(pv)) + bce(pi, torch.ones_like(pi))
            lg.backward()
            if world > 1:
                allreduce_rest()
            opt_g.step()
            losses.update(dis_img=li, dis_vid=lv, gen=lg)

        w0 = [p.detach().clone() for p in ode_params]
        for _ in range(warmup):
            iteration()
        if on_gpu:
            torch.cuda.synchronize()
        ode_ev.clear()
        import time
        if world > 1:
            dist.barrier()
        if on_gpu:
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
        t0 = time.perf_counter()
        for _ in range(iters):
            iteration()
        if on_gpu:
            e.record()
            torch.cuda.synchronize()
            total = s.elapsed_time(e)
            ode_fwd = sum(a.elapsed_time(b) for a, b in ode_ev)
        else:
            total, ode_fwd = (time.perf_counter() - t0) * 1e3, float("nan")
        tt = torch.tensor([total], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        moved = max(float((p.detach() - q).abs().max()) for p, q in zip(ode_params, w0))
        if rank != 0:
            return None
        return dict(workload="configs[3]: DP MoCoGAN-ODE step, synthetic (B,3,16,64,64) clips, D=64/H=256 motion ODE (rk4 + adjoint)",
                    nets=which, n_gpus=world, videos_per_gpu=B, iters=iters, precision=precision,
                    iters_per_s=iters / float(tt.item()) * 1e3, ms_per_iter=float(tt.item()) / iters,
                    ode_forward_ms_per_iter=ode_fwd / iters, ode_trajectories_per_iter=3 * (B + B * T * 2),
                    ode_forward_share=ode_fwd / total, ode_param_max_update=moved,
                    losses={k: float(v.detach()) for k, v in losses.items()})
    finally:
        if solver_module is not None:
            if prev_mod is None:
                sys.modules.pop("torchdiffeq", None)
            else:
                sys.modules["torchdiffeq"] = prev_mod


def main():
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--batch", type=int, default=32, help="videos per GPU (reference: batch_size = 32)")
    ap.add_argument("--fp32", action="store_true", help="FP32 wide-field kernels instead of the tcgen05 bf16 path")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    out_fd = 1
    if world > 1:
        sys.stdout.flush()
        out_fd = os.dup(1)
        os.dup2(2, 1)
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = run_harness(args.iters, args.batch, "fp32" if args.fp32 else "bf16")
    if res is not None:
        os.write(out_fd, (json.dumps(res) + "\n").encode())
    if world > 1:
        dist.barrier()
        os._exit(0)


if __name__ == "__main__":
    main()
