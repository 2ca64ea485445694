"""Synthetic "Blender-shaped" scene used by the benchmark and the parity tests.

800x800 images, camera_angle_x = 0.6911112 (fx = fy = 1111.11, cx = cy = 400, provider.py:271-281), 100 cameras on the
upper hemisphere at radius 4.0311 looking at the origin, rays from the reference's get_rays maths
(reconstruction/nerf/utils.py:136-147: pixel centre + 0.5, normalised directions, camera looks along +z),
occupancy = ball |p| <= radius in every cascade, random-init parameters.  Everything is generated from a seeded
torch.Generator on the CPU so that every rank / every implementation sees identical inputs.
"""
import math
from dataclasses import dataclass

import numpy as np
import torch

CONFIGS = {
    # name: channels, final resolution, wavelet upscale S (levels = log2 S), hidden, rays
    "tiny": dict(C=16, R=128, S=2, hidden=64, rays=2048),            # CI / smoke
    "cpu": dict(C=16, R=512, S=8, hidden=64, rays=4096),             # BASELINE.json configs[0] (65536 points)
    "small": dict(C=16, R=1024, S=16, hidden=64, rays=60000),        # configs[1], final stage
    "base_light": dict(C=32, R=2048, S=32, hidden=64, rays=60000),   # configs[2] -- the headline
    "large": dict(C=48, R=2048, S=32, hidden=128, rays=60000),       # configs[3]
}

W_IMG = H_IMG = 800
CAMERA_ANGLE_X = 0.6911112
RADIUS = 4.0311


@dataclass
class Scene:
    poses: torch.Tensor       # [100, 4, 4] cam2world
    intrinsics: tuple         # fx, fy, cx, cy
    bound: float = 1.5
    min_near: float = 0.2


def make_poses(n=100, seed=0):
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(n, generator=g)
    phi = torch.rand(n, generator=g) * 2 * math.pi
    z = u * 0.9 + 0.05                                  # upper hemisphere, away from the pole / equator
    r = torch.sqrt(1 - z * z)
    pos = torch.stack([r * torch.cos(phi), r * torch.sin(phi), z], dim=-1) * RADIUS
    fwd = -pos / pos.norm(dim=-1, keepdim=True)         # camera +z looks at the origin
    up = torch.tensor([0.0, 0.0, 1.0]).expand_as(fwd)
    right = torch.cross(fwd, up, dim=-1)
    right = right / right.norm(dim=-1, keepdim=True)
    down = torch.cross(fwd, right, dim=-1)
    poses = torch.eye(4).repeat(n, 1, 1)
    poses[:, :3, 0], poses[:, :3, 1], poses[:, :3, 2], poses[:, :3, 3] = right, down, fwd, pos
    return poses


def make_scene(seed=0):
    f = 0.5 * W_IMG / math.tan(0.5 * CAMERA_ANGLE_X)
    return Scene(poses=make_poses(100, seed), intrinsics=(f, f, W_IMG / 2, H_IMG / 2))


def rays_for_pixels(scene, img_idx, pix_idx):
    """get_rays maths (nerf/utils.py:136-147) for explicit (image, pixel) pairs -> rays_o, rays_d [N,3] (CPU fp32)."""
    fx, fy, cx, cy = scene.intrinsics
    i = (pix_idx % W_IMG).float() + 0.5
    j = (pix_idx // W_IMG).float() + 0.5
    zs = torch.ones_like(i)
    d = torch.stack(((i - cx) / fx * zs, (j - cy) / fy * zs, zs), dim=-1)
    d = d / torch.norm(d, dim=-1, keepdim=True)
    R = scene.poses[img_idx, :3, :3]
    rays_d = torch.einsum('nk,njk->nj', d, R)           # d @ R^T
    rays_o = scene.poses[img_idx, :3, 3]
    return rays_o.contiguous(), rays_d.contiguous()


def sample_batch(scene, n_rays, gen):
    """Uniform (image, pixel) pairs with replacement + random target colours; mirrors shuffle_data/select_batch."""
    img = torch.randint(0, scene.poses.shape[0], (n_rays,), generator=gen)
    pix = torch.randint(0, W_IMG * H_IMG, (n_rays,), generator=gen)
    rays_o, rays_d = rays_for_pixels(scene, img, pix)
    target = torch.rand(n_rays, 3, generator=gen)
    return rays_o, rays_d, target


def full_frame(scene, img_idx=0):
    pix = torch.arange(W_IMG * H_IMG)
    img = torch.full_like(pix, img_idx)
    return rays_for_pixels(scene, img, pix)


def _morton_np(x, y, z):
    def spread(v):
        v = v.astype(np.uint64)
        v = (v * 0x00010001) & 0xFF0000FF
        v = (v * 0x00000101) & 0x0F00F00F
        v = (v * 0x00000011) & 0xC30C30C3
        v = (v * 0x00000005) & 0x49249249
        return v & 0xFFFFFFFF
    return (spread(x) | (spread(y) << 1) | (spread(z) << 2)).astype(np.int64)


def ball_density_grid(bound=1.5, radius=0.75, value=1.0, H=128):
    """density_grid [cascade, H^3] (Morton order): `value` where the cell centre (renderer.py:478-481) lies in the ball."""
    cascade = 1 + math.ceil(math.log2(bound))
    a = np.arange(H)
    xx, yy, zz = np.meshgrid(a, a, a, indexing='ij')
    idx = _morton_np(xx.reshape(-1), yy.reshape(-1), zz.reshape(-1))
    grid = np.zeros((cascade, H ** 3), np.float32)
    base = np.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], -1).astype(np.float32)
    base = 2 * base / (H - 1) - 1
    for cas in range(cascade):
        b = min(2 ** cas, bound)
        p = base * (b - b / H)
        inside = (p * p).sum(-1) <= radius * radius
        grid[cas, idx[inside]] = value
    return torch.from_numpy(grid)


def packbits_cpu(grid, thresh):
    bits = (grid.reshape(-1).numpy() > thresh)
    return torch.from_numpy(np.packbits(bits, bitorder='little'))


def init_model_(model, seed=0, coef_sigma=0.05):
    """Random-init parameters exactly as SURVEY.md 8d prescribes (seeded, generated on CPU, copied to the device)."""
    g = torch.Generator().manual_seed(seed)
    enc = model.encoder
    with torch.no_grad():
        enc.planes_features.copy_(0.1 * torch.randn(enc.planes_features.shape, generator=g))
        for p in enc.planes_features_wavelet_coefs:
            p.copy_((coef_sigma * torch.randn(p.shape, generator=g)).to(p.device))
        for lin in list(model.sigma_net) + list(model.color_net):
            o, i = lin.weight.shape
            lin.weight.copy_(((torch.rand(o, i, generator=g) * 2 - 1) / math.sqrt(i)).to(lin.weight.device))
    enc.reset_cahce()
    return model


def install_ball_occupancy(model, radius=0.75):
    grid = ball_density_grid(model.bound, radius, value=1.0, H=model.grid_size)
    model.density_grid.copy_(grid.to(model.density_grid.device))
    model.density_bitfield.copy_(packbits_cpu(grid, 0.5).to(model.density_bitfield.device))
    model.mean_density = float(grid.clamp(min=0).mean())
    model.iter_density = 16
    model.mark_bitfield_changed()
    return model
