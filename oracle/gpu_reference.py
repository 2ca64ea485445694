"""TEST / BENCH INFRASTRUCTURE ONLY (oracle) -- the GPU reference baseline of BASELINE.md section 4 (last bullet):
one training step of the reference's hot path on a CUDA device, assembled from

  * the reference's OWN CUDA kernels, compiled unmodified (oracle/_ref/_raymarching_ref.so, _shencoder_ref.so; built by
    oracle/build_ref.py from /root/reference/aux_libs/*/src): near_far_from_aabb, march_rays_train,
    composite_rays_train_forward/backward, sh_encode_forward -- called with the argument lists of the reference's
    aux_libs/raymarching/raymarching.py:19-291 and aux_libs/shencoder/sphere_harmonics.py:14-58 (zero-filled
    [M,*] sample buffers per step included);
  * the torch library ops the reference's Python reaches: the pytorch_wavelets IDWT restated with the same
    conv_transpose2d calls (oracle/wavelet.py; the package itself is not installable here), F.grid_sample,
    autocast nn.Linear / relu / sigmoid / trunc_exp (oracle/field.py), GradScaler;

in the order of Trainer.train_one_epoch2 / train_step (reconstruction/nerf/utils.py:1138-1166, 532-679) and
NeRFRenderer.run_cuda (reconstruction/nerf/renderer.py:269-321).  It says what the hand-written kernels of the product buy
over the unfused reference data flow on the SAME GPU.  Never imported by the product package.
"""
import torch

from . import build_ref
from . import field as of
from . import wavelet as ow


def _kernels():
    rm, sh = build_ref.load_ref("raymarching"), build_ref.load_ref("shencoder")
    if rm is None or sh is None:
        raise RuntimeError("oracle/_ref/*.so not built (python oracle/build_ref.py all)")
    return rm, sh


class _Composite(torch.autograd.Function):
    """aux_libs/raymarching/raymarching.py:238-291 over the compiled reference kernels."""

    @staticmethod
    def forward(ctx, rm, sigmas, rgbs, deltas, rays, T_thresh):
        sigmas, rgbs = sigmas.float().contiguous(), rgbs.float().contiguous()
        M, N = sigmas.shape[0], rays.shape[0]
        ws = torch.empty(N, device=sigmas.device)
        depth = torch.empty(N, device=sigmas.device)
        image = torch.empty(N, 3, device=sigmas.device)
        rm.composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, T_thresh, ws, depth, image)
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, ws, image)
        ctx.aux = (rm, M, N, T_thresh)
        return ws, depth, image

    @staticmethod
    def backward(ctx, g_ws, g_depth, g_image):
        sigmas, rgbs, deltas, rays, ws, image = ctx.saved_tensors
        rm, M, N, T_thresh = ctx.aux
        gs, gc = torch.zeros_like(sigmas), torch.zeros_like(rgbs)
        rm.composite_rays_train_backward(g_ws.float().contiguous(), g_image.float().contiguous(), sigmas, rgbs, deltas, rays, ws, image,
                                         M, N, T_thresh, gs, gc)
        return None, gs, gc, None, None, None


def training_step(rm, sh, pf, coefs, weights, bitfield, batch, mean_count, scaler, bound=1.5, min_near=0.2, max_steps=1024, lam=0.2):
    rays_o, rays_d, target = batch
    N = rays_o.shape[0]
    dev = rays_o.device
    planes = ow.build_planes(pf, coefs)                                   # outside autocast: fp32 (utils.py:1138-1140)
    with torch.autocast("cuda", dtype=torch.float16):
        aabb = torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32, device=dev)
        nears, fars = torch.empty(N, device=dev), torch.empty(N, device=dev)
        rm.near_far_from_aabb(rays_o, rays_d, aabb, N, min_near, nears, fars)
        M = mean_count + 128 - mean_count % 128
        xyzs, dirs, deltas = torch.zeros(M, 3, device=dev), torch.zeros(M, 3, device=dev), torch.zeros(M, 2, device=dev)
        rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
        counter = torch.zeros(2, dtype=torch.int32, device=dev)
        noises = torch.rand(N, device=dev)
        rm.march_rays_train(rays_o, rays_d, bitfield, bound, 0.0, max_steps, N, 2, 128, M, nears, fars, xyzs, dirs, deltas, rays, counter, noises)
        feat = of.sample_planes(planes, xyzs, bound)                      # projection + F.grid_sample (autocast-fp32 op)
        W1, W2, W3, W4, W5 = weights
        h = torch.relu(torch.nn.functional.linear(feat, W1))
        h = torch.nn.functional.linear(h, W2)
        sigma = of.trunc_exp(h[:, 0].float())
        shd = torch.empty(M, 16, device=dev)
        sh.sh_encode_forward(dirs, shd, M, 3, 4, None)
        h = torch.cat([shd, h[:, 1:]], dim=-1)
        h = torch.relu(torch.nn.functional.linear(h, W3))
        h = torch.relu(torch.nn.functional.linear(h, W4))
        rgb = torch.sigmoid(torch.nn.functional.linear(h, W5))
        ws, depth, image = _Composite.apply(rm, sigma, rgb, deltas, rays, 1e-4)
        image = image + (1 - ws).unsqueeze(-1) * 0.0
        loss = ((image - target) ** 2).mean(-1).mean()
        total = sum(v.numel() for v in coefs)
        loss = loss + lam * sum(v.abs().mean() * (v.numel() / total) for v in coefs) / len(coefs)
        scaler.scale(loss).backward()
    return loss.detach(), counter


def timed_training_steps(C, R, S, hidden, batches, mean_count, radius=0.75, steps=3, seed=0):
    """-> dict(ms_per_step, rays_per_s, M, note): `steps` timed fwd+bwd steps (after one warm-up) on the caller's device."""
    import math
    from trinerflet_b200 import scene
    rm, sh = _kernels()
    dev = batches[0][0].device
    g = torch.Generator().manual_seed(seed)
    L = int(round(math.log2(S)))
    n0 = R // S
    pf = (0.1 * torch.randn(3, C, n0, n0, generator=g)).to(dev).requires_grad_(True)
    coefs = [(0.05 * torch.randn(3, C, 3, n0 * 2 ** l, n0 * 2 ** l, generator=g)).to(dev).requires_grad_(True) for l in range(L)]
    weights = [w.to(dev).requires_grad_(True) for w in of.init_mlp_weights(C, hidden, hidden, gen=g)]
    bitfield = scene.packbits_cpu(scene.ball_density_grid(1.5, radius), 0.5).to(dev)
    scaler = torch.amp.GradScaler("cuda")
    params = [pf] + coefs + weights

    def one(b):
        for p in params:
            p.grad = None
        return training_step(rm, sh, pf, coefs, weights, bitfield, b, mean_count, scaler)

    loss, counter = one(batches[0])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss, counter = one(batches[(i + 1) % len(batches)])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    N = batches[0][0].shape[0]
    return {"ms_per_step": ms, "rays_per_s": N / (ms * 1e-3), "M": int(counter[0]), "loss": float(loss), "steps": steps,
            "note": "reference CUDA kernels (oracle/_ref, unmodified) + torch library ops (conv_transpose2d IDWT restating pytorch_wavelets, "
                    "grid_sample, autocast nn.Linear), fwd+bwd, eager, same rays / occupancy / sizes; a comparator, not the product"}
