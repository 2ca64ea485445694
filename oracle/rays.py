"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference's step feeder.

  get_rays_np          reconstruction/nerf/utils.py:64-149 (the N=-1 / explicit-inds branch: :80-82 pixel centres,
                       :136-147 directions, normalisation, rotation, origin broadcast)
  shuffled_batch_np    shuffle_data + select_batch (utils.py:228-243) and the ray tables of NeRFDataset_all
                       (nerf/provider.py:683-711): rows perm[b*bs:(b+1)*bs] of the flattened [B*H*W] tables

Pinned by tests/golden/rays_ref.npz, the outputs of the reference's own get_rays / shuffle_data / select_batch imported
from /root/reference on the CPU (tests/golden/make_rays_golden.py): bit-equal.  numpy fp32, one rounding per operation;
the two reductions (squared norm, 3-term dot products) use fused multiply-adds in index order, which is what ATen's CPU
kernels do for torch.norm / matmul here (found by matching the golden vectors bit for bit).
"""
import numpy as np


def _fma(a, b, c):
    """fp32 fused multiply-add: the product of two fp32 numbers is exact in fp64; one rounding to fp64 of the sum and one
    to fp32 (double rounding differs from a true fma only in ~2^-29 of the cases)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


def pixel_dirs_np(intrinsics, W, pix):
    fx, fy, cx, cy = (np.float32(v) for v in intrinsics)
    pix = np.asarray(pix, np.int64)
    i = (pix % W).astype(np.float32) + np.float32(0.5)
    j = (pix // W).astype(np.float32) + np.float32(0.5)
    xs = (i - cx) / fx
    ys = (j - cy) / fy
    zs = np.ones_like(xs)
    nrm = np.sqrt(_fma(zs, zs, _fma(ys, ys, xs * xs)))
    return np.stack([xs / nrm, ys / nrm, zs / nrm], -1).astype(np.float32)


def get_rays_np(poses, intrinsics, H, W, inds=None):
    """poses [B,4,4] fp32; inds None (all H*W pixels) or int [N] / [B,N] -> rays_o, rays_d [B,N,3], inds [B,N]"""
    poses = np.asarray(poses, np.float32)
    B = poses.shape[0]
    if inds is None:
        inds = np.arange(H * W, dtype=np.int64)
    inds = np.broadcast_to(np.asarray(inds, np.int64), (B, np.asarray(inds).shape[-1]))
    d = pixel_dirs_np(intrinsics, W, inds)                                    # [B,N,3]
    R = poses[:, :3, :3]
    rays_d = np.zeros_like(d)
    for r in range(3):                                                        # dir @ R^T, fma chain in index order
        rays_d[..., r] = _fma(d[..., 2], R[:, None, r, 2], _fma(d[..., 1], R[:, None, r, 1], d[..., 0] * R[:, None, r, 0]))
    rays_o = np.broadcast_to(poses[:, None, :3, 3], rays_d.shape).copy()
    return rays_o, rays_d, inds


def rays_from_ids_np(poses, intrinsics, H, W, ids, images=None):
    """flat ids into the [B*H*W] ray table -> rays_o, rays_d [n,3] (and targets [n,C] if images [B,H*W,C] given)"""
    poses = np.asarray(poses, np.float32)
    ids = np.asarray(ids, np.int64)
    img, pix = ids // (H * W), ids % (H * W)
    d = pixel_dirs_np(intrinsics, W, pix)
    R = poses[img, :3, :3]
    rays_d = np.stack([_fma(d[:, 2], R[:, r, 2], _fma(d[:, 1], R[:, r, 1], d[:, 0] * R[:, r, 0])) for r in range(3)], -1)
    rays_o = poses[img, :3, 3].copy()
    if images is None:
        return rays_o, rays_d.astype(np.float32)
    return rays_o, rays_d.astype(np.float32), np.asarray(images).reshape(-1, images.shape[-1])[ids]


def shuffled_batch_np(poses, intrinsics, H, W, images, perm, batch_idx, batch_size):
    ids = np.asarray(perm, np.int64)[batch_idx * batch_size:(batch_idx + 1) * batch_size]
    return rays_from_ids_np(poses, intrinsics, H, W, ids, images)
