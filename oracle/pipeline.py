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


def timed_components(C, R, S, hidden, n_rays_full, scene_batch, bitfield, planes_sub, r_div, n_small, n_large, seed=0):
    """Bounded-sample timing of one reference training step (fwd+bwd) on the CPU, extrapolated to the full workload:
      (a) multilevel IDWT forward+backward on planes_sub of the 3 planes (batch entries are independent: exact x3/planes_sub)
          at resolution R/r_div with the same number of levels (cost is proportional to the pixel count: x r_div^2);
      (b) march + sample + MLP + composite + loss, forward+backward down to the plane gradients, on n_small and
          n_large rays against full-size planes; a linear fit t = a + b*n gives the time at n_rays_full.
    Returns dict(step_seconds, idwt_seconds, rays_seconds, detail...)."""
    g = torch.Generator().manual_seed(seed)
    L = int(round(np.log2(S)))
    n0 = R // S
    n0 = max(n0 // r_div, 8)
    pf = (0.1 * torch.randn(planes_sub, C, n0, n0, generator=g)).requires_grad_(True)
    coefs = [(0.05 * torch.randn(planes_sub, C, 3, n0 * 2 ** l, n0 * 2 ** l, generator=g)).requires_grad_(True) for l in range(L)]
    t0 = time.perf_counter()
    planes = ow.build_planes(pf, coefs)
    planes.backward(torch.ones_like(planes))
    t_idwt_sub = time.perf_counter() - t0
    t_idwt = t_idwt_sub * (3.0 / planes_sub) * (R // S / n0) ** 2
    del planes, pf, coefs
    planes = torch.randn(3, C, R, R, generator=g).mul_(0.1).requires_grad_(True)
    weights = [w.requires_grad_(True) for w in of.init_mlp_weights(C, hidden, hidden, gen=g)]
    rays_o, rays_d, target = scene_batch
    times, Ms = [], []
    for n in (n_small, n_large):
        noises = torch.rand(n, generator=g).numpy()
        planes.grad = None
        t0 = time.perf_counter()
        image, ws, depth, M = render_train(planes, weights, rays_o[:n], rays_d[:n], bitfield, noises)
        loss = ((image - target[:n]) ** 2).mean(-1).mean()
        loss.backward()
        times.append(time.perf_counter() - t0)
        Ms.append(M)
    b = (times[1] - times[0]) / max(n_large - n_small, 1)
    a = max(times[0] - b * n_small, 0.0)
    t_rays = a + b * n_rays_full
    return dict(step_seconds=t_idwt + t_rays, idwt_seconds=t_idwt, rays_seconds=t_rays, idwt_sub_seconds=t_idwt_sub,
                planes_sub=planes_sub, r_div=R // S // n0, n_small=n_small, n_large=n_large, t_small=times[0], t_large=times[1], M_small=Ms[0], M_large=Ms[1])
