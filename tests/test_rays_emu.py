"""CPU: the step feeder.  (1) the oracle (oracle/rays.py) against the golden vectors produced by the
reference's own get_rays / shuffle_data / select_batch (tests/golden/make_rays_golden.py): bit-equal; (2) the kernel's
per-ray code (trinerflet_b200/csrc/rays_core.cuh) run by the host emulator (tests/emu/rays_emu.cpp) against the same
vectors: bit-equal, including the clamping of out-of-range ids, the implicit-range (full frame) mode and the target
gather.  The GPU test (tests/test_gpu_x_feeder.py) checks the real kernel through the C ABI."""
import ctypes
import os

import numpy as np
import pytest

from oracle import rays as R
from tests.util import build_emu


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "rays_ref.npz"))


@pytest.fixture(scope="module")
def emu():
    return build_emu("rays_emu", deps=["trinerflet_b200/csrc/rays_core.cuh"], flags=["-ffp-contract=off"])


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def emu_rays(emu, poses, intr, H, W, ids=None, first_id=0, n=None, images=None):
    poses = np.ascontiguousarray(poses, np.float32)
    if ids is not None:
        ids = np.ascontiguousarray(ids, np.int64)
        n = len(ids)
    ro, rd = np.full((n, 3), np.nan, np.float32), np.full((n, 3), np.nan, np.float32)
    ci = 0 if images is None else images.shape[-1]
    gt = None if images is None else np.full((n, ci), np.nan, np.float32)
    if images is not None:
        images = np.ascontiguousarray(images, np.float32)
    f = ctypes.c_float
    emu.emu_rays_from_ids(_p(poses), poses.shape[0], f(intr[0]), f(intr[1]), f(intr[2]), f(intr[3]), H, W, _p(ids),
                          ctypes.c_int64(first_id), n, _p(images), ci, _p(ro), _p(rd), _p(gt))
    return ro, rd, gt


def test_oracle_matches_reference_get_rays(gold):
    o, d, inds = R.get_rays_np(gold["A_poses"], gold["A_intr"], 800, 800, gold["A_inds"])
    assert np.array_equal(o, gold["A_rays_o"]) and np.array_equal(d, gold["A_rays_d"])
    o, d, inds = R.get_rays_np(gold["B_poses"], gold["B_intr"], 37, 53, gold["B_inds"])
    assert np.array_equal(o, gold["B_rays_o"]) and np.array_equal(d, gold["B_rays_d"])
    assert np.array_equal(inds, gold["B_inds"])
    o, d, inds = R.get_rays_np(gold["B_poses"], gold["B_intr"], 37, 53, None)           # N = -1: every pixel
    assert np.array_equal(o, gold["B_full_o"]) and np.array_equal(d, gold["B_full_d"])
    assert np.allclose(np.linalg.norm(d.astype(np.float64), axis=-1), 1.0, atol=2e-7)


def test_oracle_matches_reference_shuffle_select(gold):
    H, W, bs = 37, 53, int(gold["C_bs"])
    for b in (0, 3, 7):                                                                 # 7 = ragged last batch (844 rays)
        o, d, im = R.shuffled_batch_np(gold["B_poses"], gold["B_intr"], H, W, gold["C_images"], gold["C_perm"], b, bs)
        assert np.array_equal(o, gold[f"C_b{b}_rays_o"]) and np.array_equal(d, gold[f"C_b{b}_rays_d"])
        assert np.array_equal(im, gold[f"C_b{b}_images"])
    assert len(gold["C_b7_rays_o"]) == 4 * H * W - 7 * bs


def test_emulated_kernel_matches_reference(gold, emu):
    # explicit ids, Blender-shaped geometry (image b, the picked pixels)
    for b in range(5):
        ids = b * 800 * 800 + gold["A_inds"].astype(np.int64)
        ro, rd, _ = emu_rays(emu, gold["A_poses"], gold["A_intr"], 800, 800, ids)
        assert np.array_equal(ro, gold["A_rays_o"][b]) and np.array_equal(rd, gold["A_rays_d"][b])
    # implicit contiguous range = full frames of all 4 images, non-square, fx != fy
    H, W = 37, 53
    ro, rd, _ = emu_rays(emu, gold["B_poses"], gold["B_intr"], H, W, None, 0, 4 * H * W)
    assert np.array_equal(ro.reshape(4, -1, 3), gold["B_full_o"]) and np.array_equal(rd.reshape(4, -1, 3), gold["B_full_d"])
    # a range that starts inside image 2
    ro, rd, _ = emu_rays(emu, gold["B_poses"], gold["B_intr"], H, W, None, 2 * H * W + 100, 500)
    assert np.array_equal(rd, gold["B_full_d"][2, 100:600])
    # the feeder: slices of the reference's permutation, with the target gather (4 channels)
    bs = int(gold["C_bs"])
    for b in (0, 3, 7):
        ids = gold["C_perm"][b * bs:(b + 1) * bs]
        ro, rd, gt = emu_rays(emu, gold["B_poses"], gold["B_intr"], H, W, ids, images=gold["C_images"])
        assert np.array_equal(ro, gold[f"C_b{b}_rays_o"]) and np.array_equal(rd, gold[f"C_b{b}_rays_d"])
        assert np.array_equal(gt, gold[f"C_b{b}_images"])


def test_emulated_kernel_edge_cases(gold, emu):
    H, W = 37, 53
    n_total = 4 * H * W
    ids = np.array([-5, 0, n_total - 1, n_total, n_total + 123456789], np.int64)        # clamped to the table
    ro, rd, gt = emu_rays(emu, gold["B_poses"], gold["B_intr"], H, W, ids, images=gold["C_images"])
    full_d = gold["B_full_d"].reshape(-1, 3)
    assert np.array_equal(rd, full_d[[0, 0, n_total - 1, n_total - 1, n_total - 1]])
    assert np.array_equal(gt, gold["C_images"].reshape(-1, 4)[[0, 0, n_total - 1, n_total - 1, n_total - 1]])
    ro, rd, _ = emu_rays(emu, gold["B_poses"], gold["B_intr"], H, W, np.zeros(0, np.int64))   # empty batch
    assert ro.shape == (0, 3)
    # duplicates (randint sampling "may duplicate", utils.py:115) give identical rows
    ro, rd, _ = emu_rays(emu, gold["B_poses"], gold["B_intr"], H, W, np.array([77, 77, 77], np.int64))
    assert np.array_equal(rd[0], rd[1]) and np.array_equal(rd[1], rd[2])


def test_abi_argument_errors_without_gpu():
    from trinerflet_b200 import _lib
    lib = _lib.load()
    z = ctypes.c_void_p(0)
    p16 = ctypes.c_void_p(64)
    assert lib.tnl_rays_from_ids(z, 1, 1.0, 1.0, 0.0, 0.0, 4, 4, z, 0, 0, z, 0, z, z, z, None) == 0       # n = 0: nothing to do
    assert lib.tnl_rays_from_ids(z, 1, 1.0, 1.0, 0.0, 0.0, 4, 4, z, 0, 5, z, 0, z, z, z, None) == -1      # null pointers
    assert lib.tnl_rays_from_ids(p16, 1, 0.0, 1.0, 0.0, 0.0, 4, 4, z, 0, 5, z, 0, p16, p16, z, None) == -1  # fx = 0
    assert b"focal" in lib.tnl_last_error()
    assert lib.tnl_rays_from_ids(p16, 1, 1.0, 1.0, 0.0, 0.0, 4, 4, z, 10, 8, z, 0, p16, p16, z, None) == -1  # range past table
    assert lib.tnl_rays_from_ids(p16, 1, 1.0, 1.0, 0.0, 0.0, 4, 4, z, 0, 8, p16, 5, p16, p16, p16, None) == -1  # 5 channels
    assert lib.tnl_rays_from_ids(p16, 1, 1.0, 1.0, 0.0, 0.0, 4, 4, z, 0, 8, p16, 3, p16, p16, z, None) == -1  # images w/o targets
