"""CPU: known-answer tests for the plain-C ray-marching oracle (oracle/raymarch.c) and, when the fixture generated
from the UNMODIFIED reference kernels on a B200 is present (tests/golden/raymarch_ref.npz), bit-exact agreement with it."""
import os

import numpy as np
import pytest

from oracle import raymarch as orc
from tests.util import random_bitfield, synthetic_rays

BOUND, CAS, H = 1.5, 2, 128
AABB = np.array([-BOUND] * 3 + [BOUND] * 3, np.float32)


def test_morton_roundtrip_and_packbits():
    a = np.arange(128, dtype=np.int32)
    c = np.stack(np.meshgrid(a[::3], a[::5], a[::7], indexing='ij'), -1).reshape(-1, 3).astype(np.int32)
    idx = orc.morton3D(c)
    assert idx.min() >= 0 and idx.max() < 128 ** 3 and len(np.unique(idx)) == len(idx)
    assert np.array_equal(orc.morton3D_invert(idx), c)
    assert orc.morton3D(np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [127, 127, 127]], np.int32)).tolist() == [1, 2, 4, 128 ** 3 - 1]
    grid = np.random.default_rng(0).random(4096).astype(np.float32)
    assert np.array_equal(orc.packbits(grid, 0.4), np.packbits(grid > 0.4, bitorder='little'))


def test_near_far():
    o = np.array([[0, 0, -4], [0, 0, -4], [0, 0, -1], [5, 5, 5]], np.float32)
    d = np.array([[0, 0, 1], [0, 1, 0], [0, 0, 1], [1, 0, 0]], np.float32)
    with np.errstate(divide='ignore'):
        n, f = orc.near_far_from_aabb(o, d, AABB, 0.2)
    assert n[0] == 2.5 and f[0] == 5.5
    assert n[1] == np.finfo(np.float32).max and f[1] == np.finfo(np.float32).max      # miss
    assert n[2] == np.float32(0.2) and f[2] == 2.5                                     # origin inside: near clamped
    assert n[3] == np.finfo(np.float32).max


def test_march_empty_and_full_grid():
    o, d = synthetic_rays(256, 1)
    nears, fars = orc.near_far_from_aabb(o, d, AABB, 0.2)
    zeros = np.zeros(256, np.float32)
    empty = np.zeros(CAS * H ** 3 // 8, np.uint8)
    x, dd, dl, rays, cnt = orc.march_rays_train(o, d, BOUND, empty, CAS, H, nears, fars, zeros, 1000)
    assert cnt.tolist() == [0, 256] and rays[:, 2].sum() == 0 and np.abs(x).sum() == 0
    full = np.full(CAS * H ** 3 // 8, 255, np.uint8)
    x, dd, dl, rays, cnt = orc.march_rays_train(o, d, BOUND, full, CAS, H, nears, fars, zeros, 256 * 1024)
    dt_min = np.float32(2 * np.float32(1.7320508075688772) / np.float32(1024))
    for n in range(0, 256, 17):
        t, k = nears[n], 0
        while t < fars[n] and k < 1024:      # fp32 running sum t0 + dt + dt + ...
            t = np.float32(t + dt_min); k += 1
        assert rays[n, 2] == k
        assert np.array_equal(rays[n], [n, rays[:n, 2].sum(), k])
    seg = dl[rays[5, 1]:rays[5, 1] + rays[5, 2]]
    assert np.all(seg[:, 0] == dt_min) and np.abs(seg[:, 1] - dt_min).max() < 1e-6


def test_composite_against_numpy():
    rng = np.random.default_rng(0)
    counts = np.array([0, 5, 40, 1, 300], np.int32)
    offs = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int32)
    rays = np.stack([np.array([3, 1, 4, 0, 2], np.int32), offs, counts], -1)
    M = int(counts.sum())
    sig = (rng.random(M) * 30).astype(np.float32); rgb = rng.random((M, 3)).astype(np.float32)
    deltas = np.stack([np.full(M, 0.0034, np.float32), np.full(M, 0.0034, np.float32)], -1)
    ws, dp, im = orc.composite_rays_train_forward(sig, rgb, deltas, rays, 1e-4)
    for idx, off, k in rays:
        a = 1 - np.exp(-sig[off:off + k].astype(np.float64) * 0.0034)
        T = np.concatenate([[1.0], np.cumprod(1 - a)])[:-1]
        stop = np.nonzero(np.cumprod(1 - a) < 1e-4)[0]
        e = (stop[0] + 1) if len(stop) else k
        w = (a * T)[:e]
        assert abs(ws[idx] - w.sum()) < 1e-5 and np.abs(im[idx] - (w[:, None] * rgb[off:off + e]).sum(0)).max() < 1e-5
    gws, gim = rng.normal(size=5).astype(np.float32), rng.normal(size=(5, 3)).astype(np.float32)
    gs, gc = orc.composite_rays_train_backward(gws, gim, sig, rgb, deltas, rays, ws, im, 1e-4)
    # finite differences on one sample of the 40-sample ray (fp32 forward => loose tolerance)
    j = offs[2] + 3
    def loss(s):
        w_, _, i_ = orc.composite_rays_train_forward(s, rgb, deltas, rays, 1e-4)
        return float((w_.astype(np.float64) * gws).sum() + (i_.astype(np.float64) * gim).sum())
    sp, sm = sig.copy(), sig.copy(); sp[j] += 0.05; sm[j] -= 0.05
    fd = (loss(sp) - loss(sm)) / 0.1
    assert abs(fd - gs[j]) <= 0.05 * max(1e-3, abs(gs[j]))
    assert np.abs(gc[offs[0]:offs[0] + 0]).sum() == 0


def test_reference_kernel_fixture(golden_dir):
    """Outputs of the reference's own CUDA kernels (oracle/_ref, built unmodified from /root/reference, run on a B200)."""
    path = os.path.join(golden_dir, "raymarch_ref.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/raymarch_ref.npz not generated yet (tests/golden/make_raymarch_golden.py, needs a GPU)")
    g = np.load(path)
    o, d, bits, noises = g["rays_o"], g["rays_d"], g["bitfield"], g["noises"]
    nears, fars = orc.near_far_from_aabb(o, d, AABB, 0.2)
    assert np.array_equal(nears.view(np.uint32), g["nears"].view(np.uint32)) and np.array_equal(fars.view(np.uint32), g["fars"].view(np.uint32))
    M = int(g["M"])
    x, dd, dl, rays, cnt = orc.march_rays_train(o, d, BOUND, bits, CAS, H, nears, fars, noises, M, float(g["dt_gamma"]), int(g["max_steps"]))
    assert np.array_equal(cnt, g["counter"])
    ref_rays = g["rays"][np.argsort(g["rays"][:, 0])]
    assert np.array_equal(ref_rays[:, 2], rays[:, 2])
    for n in range(0, len(o), 7):
        a, b, k = ref_rays[n, 1], rays[n, 1], rays[n, 2]
        assert np.array_equal(g["xyzs"][a:a + k].view(np.uint32), x[b:b + k].view(np.uint32))
        assert np.array_equal(g["deltas"][a:a + k].view(np.uint32), dl[b:b + k].view(np.uint32))
    ws, dp, im = orc.composite_rays_train_forward(g["sigmas"], g["rgbs"], g["deltas"], g["rays"], 1e-4)
    assert np.abs(ws - g["weights_sum"]).max() < 2e-5 and np.abs(im - g["image"]).max() < 2e-5
    assert np.array_equal(orc.packbits(g["grid"], float(g["thresh"])), g["packed"])
    assert np.array_equal(orc.morton3D(g["coords"]), g["morton"])
