"""configs[3] end to end: a data-parallel MoCoGAN-ODE training step on synthetic UCF-shaped clips.

    python scripts/train_dp_harness.py                                     # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29521 \
        scripts/train_dp_harness.py [--iters K] [--batch B] [--fp32]

Shape of the loop = the reference's train() (ucf_moco_ode.py:53-191 / mnist_moco_ode.py:113-165): per iteration two
discriminator rounds (image + video discriminator on real clips and on generator samples drawn under no_grad) and one
generator step whose backward runs through the latent-motion ODE; three Adam optimisers (lr 2e-4, betas (0.5, 0.999),
weight decay 1e-5); BCE-with-logits.  What is NOT the reference: the conv generator / discriminators below are stand-ins of
the same tensor shapes written for this harness (they "stay PyTorch" and are wrapped in DDP), and the data are random
clips (B, 3, 16, 64, 64).  What IS this repository: `sample_z_m` is the reference's (models/mocogan_ode.py:133-148) with the
larger motion code (D=64, H=256); its `odeint_adjoint(..., method='rk4')` resolves to gan_ode_b200 through the shim —
tcgen05 forward + tcgen05 adjoint in bf16 mode — and the ODE parameter gradient is all-reduced by this library (NCCL here:
33 088 floats), not by DDP.  Reports iterations/s and the share of the step spent in the ODE solves (CUDA events).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn as nn

import gan_ode_b200 as gode

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--batch", type=int, default=32, help="videos per GPU (reference: batch_size = 32)")
ap.add_argument("--fp32", action="store_true", help="FP32 wide-field kernels instead of the tcgen05 bf16 path")
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
OUT_FD = 1
if world > 1:
    sys.stdout.flush()
    OUT_FD = os.dup(1)
    os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)
    gode.config.grad_allreduce = True
gode.install_shims()
gode.config.layout = "btd"
gode.config.precision = "fp32" if args.fp32 else "bf16"

DZM, DH, DZC, T = 64, 256, 50, 16


class ODEFunc(nn.Module):  # the reference's module interface (models/mocogan_ode.py:6-17)
    def __init__(self, dim, dim_hidden):
        super().__init__()
        self.fn = nn.Sequential(nn.Linear(dim, dim_hidden), nn.Tanh(), nn.Linear(dim_hidden, dim))

    def forward(self, t, x):
        return self.fn(x)


class Motion(nn.Module):
    """Latent-motion sampler, call for call models/mocogan_ode.py:123-148 with dim_z_motion = 64."""

    def __init__(self):
        super().__init__()
        self.ode_fn = ODEFunc(DZM, DH)
        self.linear = nn.Sequential(nn.Linear(DZM, 64), nn.LeakyReLU(0.2), nn.Linear(64, DZM), nn.LeakyReLU(0.2))

    def sample_z_m(self, n):
        from torchdiffeq import odeint_adjoint as odeint  # -> gan_ode_b200 through the shim
        x = self.linear(torch.randn(n, DZM, device=dev))
        z = odeint(self.ode_fn, x, torch.linspace(0, 1, T).float(), method='rk4')
        return z.transpose(0, 1).reshape(-1, DZM)


class FrameGenerator(nn.Module):  # stand-in: (B*T, DZC + DZM) -> (B*T, 3, 64, 64)
    def __init__(self, ngf=64):
        super().__init__()
        chans = [DZC + DZM, ngf * 8, ngf * 4, ngf * 2, ngf]
        layers = []
        for i in range(4):
            layers += [nn.ConvTranspose2d(chans[i], chans[i + 1], 4, 1 if i == 0 else 2, 0 if i == 0 else 1, bias=False),
                       nn.BatchNorm2d(chans[i + 1]), nn.ReLU(True)]
        layers += [nn.ConvTranspose2d(ngf, 3, 4, 2, 1, bias=False), nn.Tanh()]
        self.main = nn.Sequential(*layers)

    def forward(self, z):
        return self.main(z.view(z.shape[0], -1, 1, 1))


class ImageD(nn.Module):
    def __init__(self, ndf=64):
        super().__init__()
        self.main = nn.Sequential(nn.Conv2d(3, ndf, 4, 2, 1), nn.LeakyReLU(0.2), nn.Conv2d(ndf, ndf * 2, 4, 2, 1), nn.BatchNorm2d(ndf * 2),
                                  nn.LeakyReLU(0.2), nn.Conv2d(ndf * 2, ndf * 4, 4, 2, 1), nn.BatchNorm2d(ndf * 4), nn.LeakyReLU(0.2),
                                  nn.Conv2d(ndf * 4, 1, 4, 2, 1))

    def forward(self, x):
        return self.main(x).flatten(1).mean(1)


class VideoD(nn.Module):
    def __init__(self, ndf=64):
        super().__init__()
        self.main = nn.Sequential(nn.Conv3d(3, ndf, 4, (1, 2, 2), (0, 1, 1)), nn.LeakyReLU(0.2),
                                  nn.Conv3d(ndf, ndf * 2, 4, (1, 2, 2), (0, 1, 1)), nn.BatchNorm3d(ndf * 2), nn.LeakyReLU(0.2),
                                  nn.Conv3d(ndf * 2, ndf * 4, 4, (1, 2, 2), (0, 1, 1)), nn.BatchNorm3d(ndf * 4), nn.LeakyReLU(0.2),
                                  nn.Conv3d(ndf * 4, 1, 4, (1, 2, 2), (0, 1, 1)))

    def forward(self, x):
        return self.main(x).flatten(1).mean(1)


torch.manual_seed(0)
motion, framegen, dimg, dvid = Motion().to(dev), FrameGenerator().to(dev), ImageD().to(dev), VideoD().to(dev)
if world > 1:  # conv nets: DDP; ODE parameters: this library's all-reduce (they must not be in a DDP bucket as well)
    framegen = nn.parallel.DistributedDataParallel(framegen, device_ids=[local], broadcast_buffers=False)
    dimg = nn.parallel.DistributedDataParallel(dimg, device_ids=[local], broadcast_buffers=False)
    dvid = nn.parallel.DistributedDataParallel(dvid, device_ids=[local], broadcast_buffers=False)
    for p in motion.linear.parameters():
        dist.broadcast(p.data, 0)
    for p in motion.ode_fn.parameters():
        dist.broadcast(p.data, 0)
adam = dict(lr=2e-4, betas=(0.5, 0.999), weight_decay=1e-5)
opt_g = torch.optim.Adam(list(motion.parameters()) + list(framegen.parameters()), **adam)
opt_i, opt_v = torch.optim.Adam(dimg.parameters(), **adam), torch.optim.Adam(dvid.parameters(), **adam)
bce = nn.BCEWithLogitsLoss()
B = args.batch
ode_ms = [0.0]


def timed_codes(n):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    z = motion.sample_z_m(n)
    b.record()
    ode_ms.append((a, b))
    return z


def sample_videos(n):  # models/mocogan.py:259-285: content code repeated over the frames, motion code per frame
    zc = torch.randn(n, DZC, device=dev).repeat_interleave(T, 0)
    frames = framegen(torch.cat([zc, timed_codes(n)], 1))
    return frames.view(n, T, 3, 64, 64).permute(0, 2, 1, 3, 4)


def sample_images(n):  # models/mocogan.py:287-295: sample_z_video(n*T*2) trajectories, n random frames kept
    m = n * T * 2  # = 1024 trajectories for n = 32: the reference over-generates 512x
    zc = torch.randn(m, DZC, device=dev).repeat_interleave(T, 0)
    z = torch.cat([zc, timed_codes(m)], 1)
    keep = torch.randint(0, z.shape[0], (n,), device=dev)
    return framegen(z[keep])


def d_step(D, opt, real, fake):
    opt.zero_grad(set_to_none=True)
    lr_, lf_ = D(real), D(fake.detach())
    loss = bce(lr_, torch.ones_like(lr_)) + bce(lf_, torch.zeros_like(lf_))
    loss.backward()
    opt.step()
    return loss


def iteration():
    for _ in range(2):
        real = torch.randn(B, 3, T, 64, 64, device=dev)
        with torch.no_grad():
            fi, fv = sample_images(B), sample_videos(B)
        d_step(dimg, opt_i, real[:, :, 0], fi)
        d_step(dvid, opt_v, real, fv)
    opt_g.zero_grad(set_to_none=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fv, fi = sample_videos(B), sample_images(B)
    lv, li = dvid(fv), dimg(fi)
    loss = bce(lv, torch.ones_like(lv)) + bce(li, torch.ones_like(li))
    a.record()
    loss.backward()
    b.record()
    opt_g.step()
    return (a, b)


for _ in range(3):
    iteration()
torch.cuda.synchronize()
ode_ms[:] = [0.0]
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if world > 1:
    dist.barrier()
s.record()
bw = [iteration() for _ in range(args.iters)]
e.record()
torch.cuda.synchronize()
total = s.elapsed_time(e)
ode_fwd = sum(a.elapsed_time(b) for a, b in ode_ms[1:])
tt = torch.tensor([total], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    line = dict(workload="configs[3]: DP MoCoGAN-ODE step, synthetic (B,3,16,64,64) clips, D=64/H=256 motion ODE (rk4 + adjoint)",
                n_gpus=world, videos_per_gpu=B, iters=args.iters, precision=gode.config.precision,
                iters_per_s=args.iters / float(tt.item()) * 1e3, ms_per_iter=float(tt.item()) / args.iters,
                ode_forward_ms_per_iter=ode_fwd / args.iters, generator_backward_ms_per_iter=sum(a.elapsed_time(b) for a, b in bw) / args.iters,
                ode_trajectories_per_iter=3 * (B + B * T * 2), ode_forward_share=ode_fwd / total)
    os.write(OUT_FD, (json.dumps(line) + "\n").encode())
if world > 1:
    dist.barrier()
    os._exit(0)
