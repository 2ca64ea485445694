"""GPU: the on-device step feeder through the C ABI (`tnl_rays_from_ids`) against the golden vectors the
reference's own get_rays / shuffle_data / select_batch produced (tests/golden/rays_ref.npz) and against the oracle.

Tolerance: ray origins and gathered targets exact (copies); directions |err| <= 2.4e-7 (two fp32 ulps at 1.0).  The
kernel uses IEEE-rounded intrinsics in the order of the CPU reference, and the host emulation of the same code is
bit-equal to the golden vectors (tests/test_rays_emu.py), so in practice the difference is 0.

(The file sorts after the other GPU test files on purpose: it was written after this round's GPU budget was spent, its
first run on hardware is the driver's round-end run.)"""
import os

import numpy as np
import pytest
import torch

from oracle import rays as R

pytestmark = pytest.mark.gpu
TOL_D = 2.4e-7


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "rays_ref.npz"))


def _np(t):
    return t.detach().float().cpu().numpy()


def test_rays_from_ids_matches_reference_golden(gold):
    from trinerflet_b200 import rays
    poses = torch.from_numpy(gold["A_poses"]).cuda()
    for b in range(5):
        ids = torch.from_numpy(b * 800 * 800 + gold["A_inds"].astype(np.int64)).cuda()
        ro, rd = rays.rays_from_ids(poses, gold["A_intr"], 800, 800, ray_ids=ids)
        assert np.array_equal(_np(ro), gold["A_rays_o"][b])
        assert np.abs(_np(rd) - gold["A_rays_d"][b]).max() <= TOL_D
    # implicit range: every pixel of 4 small non-square images, then a range starting inside image 2
    H, W = 37, 53
    poses = torch.from_numpy(gold["B_poses"]).cuda()
    ro, rd = rays.rays_from_ids(poses, gold["B_intr"], H, W, first_id=0, n=4 * H * W)
    assert np.array_equal(_np(ro).reshape(4, -1, 3), gold["B_full_o"])
    assert np.abs(_np(rd).reshape(4, -1, 3) - gold["B_full_d"]).max() <= TOL_D
    ro, rd = rays.rays_from_ids(poses, gold["B_intr"], H, W, first_id=2 * H * W + 100, n=500)
    assert np.abs(_np(rd) - gold["B_full_d"][2, 100:600]).max() <= TOL_D
    # out-of-range ids are clamped; an empty batch is a no-op
    n_total = 4 * H * W
    ids = torch.tensor([-5, 0, n_total - 1, n_total, n_total + 123456789], dtype=torch.int64).cuda()
    ro, rd = rays.rays_from_ids(poses, gold["B_intr"], H, W, ray_ids=ids)
    full_d = gold["B_full_d"].reshape(-1, 3)
    assert np.abs(_np(rd) - full_d[[0, 0, n_total - 1, n_total - 1, n_total - 1]]).max() <= TOL_D
    ro, rd = rays.rays_from_ids(poses, gold["B_intr"], H, W, ray_ids=torch.zeros(0, dtype=torch.int64).cuda())
    assert ro.shape == (0, 3) and rd.shape == (0, 3)


def test_get_rays_drop_in_branches(gold):
    """the reference's signature and result dict; every index-selection branch (utils.py:88-134)"""
    from trinerflet_b200 import rays
    H, W = 37, 53
    intr = gold["B_intr"]
    poses = torch.from_numpy(gold["B_poses"]).cuda()

    def check(res, n):
        assert res["rays_o"].shape == (4, n, 3) and res["rays_d"].shape == (4, n, 3) and res["inds"].shape == (4, n)
        inds = res["inds"].cpu().numpy()
        assert inds.min() >= 0 and inds.max() < H * W
        o, d, _ = R.get_rays_np(gold["B_poses"], intr, H, W, inds)
        assert np.array_equal(_np(res["rays_o"]), o)
        assert np.abs(_np(res["rays_d"]) - d).max() <= TOL_D

    res = rays.get_rays(poses, intr, H, W, -1)
    check(res, H * W)
    assert torch.equal(res["inds"][3].cpu(), torch.arange(H * W))
    assert np.abs(_np(res["rays_d"]) - gold["B_full_d"]).max() <= TOL_D
    res = rays.get_rays(poses, intr, H, W, 257)                       # torch.randint, shared by the poses
    check(res, 257)
    assert torch.equal(res["inds"][0], res["inds"][1])
    res = rays.get_rays(poses, intr, H, W, 10 ** 9)                   # N is capped at H*W (utils.py:89)
    check(res, H * W)
    res = rays.get_rays(poses, intr, H, W, 160, patch_size=4)         # 10 patches of 4x4
    check(res, 160)
    p = res["inds"][0].view(10, 16).cpu()
    assert torch.equal(p - p[:, :1], (torch.arange(4).view(4, 1) * W + torch.arange(4)).view(1, 16).expand(10, 16))
    error_map = torch.rand(4, 128 * 128, generator=torch.Generator().manual_seed(1))
    res = rays.get_rays(poses, intr, H, W, 64, error_map=error_map)   # multinomial on the 128x128 error map
    check(res, 64)
    assert res["inds_coarse"].shape == (4, 64)
    # under autocast the arithmetic stays fp32 (utils.py:64 decorator)
    with torch.autocast("cuda", dtype=torch.float16):
        res = rays.get_rays(poses, intr, H, W, -1)
    assert res["rays_d"].dtype == torch.float32


def test_feeder_equals_reference_shuffle_select(gold):
    from trinerflet_b200 import rays
    H, W, bs = 37, 53, int(gold["C_bs"])
    poses = torch.from_numpy(gold["B_poses"]).cuda()
    images = torch.from_numpy(gold["C_images"]).view(4, H, W, 4).cuda()
    feeder = rays.RayFeeder(poses, gold["B_intr"], H, W, images)
    assert feeder.steps_per_epoch(bs) == 8
    feeder.shuffle(perm=torch.from_numpy(gold["C_perm"]))             # the permutation the reference's shuffle_data drew
    for b in (0, 3, 7):                                               # 7: ragged last batch
        data = feeder.select_batch(b, bs)
        assert np.array_equal(_np(data["rays_o"][0]), gold[f"C_b{b}_rays_o"])
        assert np.abs(_np(data["rays_d"][0]) - gold[f"C_b{b}_rays_d"]).max() <= TOL_D
        assert np.array_equal(_np(data["images"][0]), gold[f"C_b{b}_images"])
    # own permutation: a permutation of all ray ids, reproducible from the seed; rank shards tile the batch
    f0 = rays.RayFeeder(poses, gold["B_intr"], H, W, images, seed=3).shuffle()
    assert torch.equal(torch.sort(f0.perm).values.cpu(), torch.arange(4 * H * W))
    whole = f0.select_batch(2, 1001)
    parts = []
    for r in range(2):
        fr = rays.RayFeeder(poses, gold["B_intr"], H, W, images, seed=3, rank=r, world_size=2).shuffle()
        assert torch.equal(fr.perm, f0.perm)
        parts.append(fr.select_batch(2, 1001))
    assert parts[0]["rays_d"].shape[1] == 501 and parts[1]["rays_d"].shape[1] == 500
    for k in ("rays_o", "rays_d", "images"):
        assert torch.equal(torch.cat([parts[0][k], parts[1][k]], 1), whole[k])


def test_feeder_writes_into_given_buffers(gold):
    from trinerflet_b200 import rays
    H, W = 37, 53
    poses = torch.from_numpy(gold["B_poses"]).cuda()
    images = torch.from_numpy(gold["C_images"][..., :3].copy()).cuda()
    feeder = rays.RayFeeder(poses, gold["B_intr"], H, W, images, seed=0).shuffle()
    ref = feeder.select_batch(1, 512)
    bufs = tuple(torch.full((512, 3), float("nan"), device="cuda") for _ in range(3))
    feeder.select_batch(1, 512, out=bufs)
    for t, k in zip(bufs, ("rays_o", "rays_d", "images")):
        assert torch.equal(t, ref[k][0])
    with pytest.raises(RuntimeError):
        feeder.select_batch(1, 512, out=tuple(torch.zeros(100, 3, device="cuda") for _ in range(3)))
