"""TEST INFRASTRUCTURE ONLY (oracle) -- the reference's training-step hot path on the CPU, assembled from the
oracle pieces: C ray marching/compositing (oracle/raymarch.c) + torch wavelet reconstruction, grid_sample and
MLP heads (oracle/wavelet.py, oracle/field.py), in the order of Trainer.train_one_epoch2 / train_step
(reconstruction/nerf/utils.py:1138-1166, 532-679) and NeRFRenderer.run_cuda (reconstruction/nerf/renderer.py:257-321).
fp32 throughout (what the reference computes on a CPU: CUDA autocast is a no-op there).

Used (a) by tests as the end-to-end checker and (b) by bench.py as the timed CPU baseline / `--impl reference`
arm.  Never imported by the product package.
"""
import time

import numpy as np
import torch

from . import field as of
from . import raymarch as orc
from . import wavelet as ow


class _CompositeTrain(torch.autograd.Function):
    """aux_libs/raymarching/raymarching.py:238-291 on top of the C restatement."""

    @staticmethod
    def forward(ctx, sigmas, rgbs, deltas, rays, T_thresh):
        ws, depth, image = orc.composite_rays_train_forward(sigmas.detach().numpy(), rgbs.detach().numpy(), deltas, rays, T_thresh)
        ctx.save_for_backward(sigmas, rgbs)
        ctx.aux = (deltas, rays, ws, image, T_thresh)
        return torch.from_numpy(ws), torch.from_numpy(depth), torch.from_numpy(image)

    @staticmethod
    def backward(ctx, g_ws, g_depth, g_image):
        sigmas, rgbs = ctx.saved_tensors
        deltas, rays, ws, image, T_thresh = ctx.aux
        gs, gc = orc.composite_rays_train_backward(g_ws.contiguous().numpy(), g_image.contiguous().numpy(),
                                                   sigmas.detach().numpy(), rgbs.detach().numpy(), deltas, rays, ws, image, T_thresh)
        return torch.from_numpy(gs), torch.from_numpy(gc), None, None, None


def render_train(planes, weights, rays_o, rays_d, bitfield, noises, bound=1.5, min_near=0.2, max_steps=1024, dt_gamma=0.0,
                 bg_color=0.0, fp16=False, cascade=2, H=128, T_thresh=1e-4):
    """run_cuda training branch (renderer.py:269-321) -> image [N,3], weights_sum [N], depth [N], M."""
    aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    ro, rd = rays_o.numpy(), rays_d.numpy()
    nears, fars = orc.near_far_from_aabb(ro, rd, aabb, min_near)
    M = int(orc.march_rays_train(ro, rd, bound, bitfield, cascade, H, nears, fars, noises, 0, dt_gamma, max_steps)[4][0])
    M_alloc = M + 128 - M % 128
    xyzs, dirs, deltas, rays, _ = orc.march_rays_train(ro, rd, bound, bitfield, cascade, H, nears, fars, noises, M_alloc,
                                                       dt_gamma, max_steps)
    sigmas, rgbs = of.field_forward(planes, torch.from_numpy(xyzs), torch.from_numpy(dirs), weights, bound, fp16=fp16,
                                    recip_mul=False)
    ws, depth, image = _CompositeTrain.apply(sigmas, rgbs, deltas, rays, T_thresh)
    image = image + (1 - ws).unsqueeze(-1) * bg_color
    with np.errstate(invalid="ignore"):
        depth = torch.clamp(depth - torch.from_numpy(nears), min=0) / torch.from_numpy(fars - nears)
    return image, ws, depth, M


def wavelet_regulariser(coefs, lam):
    total = sum(v.numel() for v in coefs)
    return lam * sum(v.abs().mean() * (v.numel() / total) for v in coefs) / len(coefs)


def train_step(pf, coefs, weights, rays_o, rays_d, target, bitfield, noises, lam=0.2, loss_scale=1.0, **kw):
    """One fwd+bwd of the reference's step on the CPU; all of pf / coefs / weights must require grad. Returns loss, M.
    loss_scale: GradScaler's factor (scaler.scale(loss).backward(), nerf/utils.py:1166) -- the .grad fields then hold the
    SCALED gradients, fp16 rounding points of the emulated autocast backward included, exactly as before scaler.step()."""
    planes = ow.build_planes(pf, coefs)
    image, ws, depth, M = render_train(planes, weights, rays_o, rays_d, bitfield, noises, **kw)
    loss = ((image - target) ** 2).mean(-1).mean() + wavelet_regulariser(coefs, lam)
    (loss * loss_scale).backward()
    return float(loss), M


def timed_step_sample(C, R, S, hidden, batch, bitfield, n_sample, seed=0, planes_sub=3):
    """DIRECT timing of one reference training step (fwd+bwd) on the CPU with a bounded number of rays: the multilevel IDWT
    forward and backward run in full (all `planes_sub` of the three planes at full resolution; planes_sub < 3 only on hosts
    too small to finish, stated by the caller), the ray part (march -> grid_sample -> MLP -> composite -> MSE, forward and
    backward down to the plane gradient) runs on the first n_sample rays of `batch`.  The autograd graph is cut at the planes
    only to attribute the time (same kernels, same order as one backward call).
    Returns dict(idwt_fwd_s, rays_s, idwt_bwd_s, step_s, M)."""
    g = torch.Generator().manual_seed(seed)
    L = int(round(np.log2(S)))
    n0 = R // S
    pf = (0.1 * torch.randn(planes_sub, C, n0, n0, generator=g)).requires_grad_(True)
    coefs = [(0.05 * torch.randn(planes_sub, C, 3, n0 * 2 ** l, n0 * 2 ** l, generator=g)).requires_grad_(True) for l in range(L)]
    weights = [w.requires_grad_(True) for w in of.init_mlp_weights(C, hidden, hidden, gen=g)]
    rays_o, rays_d, target = (t[:n_sample] for t in batch)
    noises = torch.rand(n_sample, generator=g).numpy()
    t0 = time.perf_counter()
    planes = ow.build_planes(pf, coefs)
    t1 = time.perf_counter()
    full = planes if planes_sub == 3 else torch.cat([planes] * 3, 0)[:3]        # (the sampler needs three planes)
    leaf = full.detach().requires_grad_(True)
    image, ws, depth, M = render_train(leaf, weights, rays_o, rays_d, bitfield, noises)
    loss = ((image - target) ** 2).mean(-1).mean()
    loss.backward()
    t2 = time.perf_counter()
    reg = wavelet_regulariser(coefs, 0.2)
    torch.autograd.backward([planes, reg], [leaf.grad[:planes_sub], None])
    t3 = time.perf_counter()
    return dict(idwt_fwd_s=t1 - t0, rays_s=t2 - t1, idwt_bwd_s=t3 - t2, step_s=t3 - t0, M=M)


def cpu_config_benchmark(points=65536, C=16, R=512, S=8, hidden=64, iters=5, warmup=2, seed=0):
    """the reference's torch path on the CPU config -- TriPlaneVolume (C=16, R=512, S=8: three IDWT
    levels from a 64^2 base; pytorch_wavelets restated, oracle/wavelet.py) + grid_sample + sigma/color MLPs (fp32, SH restated)
    forward AND backward on `points` uniform random points in [-1.5, 1.5]^3 with random unit directions, random-init, seed 0;
    `warmup` + `iters` iterations, median.  Returns points/s, the IDWT-only GB/s (2P / t_fwd) and the breakdown."""
    g = torch.Generator().manual_seed(seed)
    L = int(round(np.log2(S)))
    n0 = R // S
    pf = (0.1 * torch.randn(3, C, n0, n0, generator=g)).requires_grad_(True)
    coefs = [(0.05 * torch.randn(3, C, 3, n0 * 2 ** l, n0 * 2 ** l, generator=g)).requires_grad_(True) for l in range(L)]
    weights = [w.requires_grad_(True) for w in of.init_mlp_weights(C, hidden, hidden, gen=g)]
    xyz = (torch.rand(points, 3, generator=g) * 2 - 1) * 1.5
    d = torch.randn(points, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    gs = torch.randn(points, generator=g)
    grgb = torch.randn(points, 3, generator=g)
    rows = []
    for it in range(warmup + iters):
        for t in [pf] + coefs + weights:
            t.grad = None
        t0 = time.perf_counter()
        planes = ow.build_planes(pf, coefs)
        t1 = time.perf_counter()
        feat = of.sample_planes(planes, xyz, 1.5)
        t2 = time.perf_counter()
        sigma, rgb, _ = of.mlp_forward(feat, d, weights)
        t3 = time.perf_counter()
        ((sigma * gs).sum() + (rgb * grgb).sum()).backward()
        t4 = time.perf_counter()
        if it >= warmup:
            rows.append((t4 - t0, t1 - t0, t2 - t1, t3 - t2, t4 - t3))
    med = [float(np.median([r[k] for r in rows])) for k in range(5)]
    P = 3 * C * R * R * 4
    return dict(points=points, points_per_s=points / med[0], fwd_bwd_seconds=med[0], idwt_fwd_seconds=med[1],
                idwt_fwd_GBps=2 * P / med[1] / 1e9, sample_fwd_seconds=med[2], mlp_fwd_seconds=med[3], backward_seconds=med[4],
                iters=iters, warmup=warmup, config=f"C={C} R={R} S={S} hidden={hidden} fp32")
