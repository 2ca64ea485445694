"""CPU: the PRODUCT kernels' source (trinerflet_b200/csrc/{raymarch,sample,sort,tiles,optim,grid,rays}.cu) executed on the
host by the CUDA-on-host shim (tests/emu/cuda_shim.h: the threads of a block are fibers, __syncthreads and the warp
collectives are real rendezvous; tests/emu/gen_kemu.py rewrites only the <<<...>>> launch sites) behind the same C ABI,
checked against the oracle exactly as the GPU tests check the device build -- at sizes the emulator finishes in seconds.

What this covers without a GPU: index arithmetic, scans, compaction, tap weights, rounding order (the fp32 intrinsics map
to correctly rounded host operations), the IDWT kernels through their real entry points (cp.async / FFMA2 have host
equivalents in idwt_core.cuh) and the mma.sync MLP kernels (the shim emulates mma.m16n8k16 and ldmatrix as warp
collectives).  What it cannot cover: the tcgen05 MLP kernels (csrc/mlp_tc.cu), `__expf` (libm here, ex2.approx on the
device), the tensor core's internal summation order, and anything about performance.  The GPU tests (-m gpu) remain the parity tests proper."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import field as of
from oracle import raymarch as orc
from tests import kemu
from tests.util import random_bitfield, synthetic_rays

BOUND, CAS, H = 1.5, 2, 128
AABB = np.array([-BOUND] * 3 + [BOUND] * 3, np.float32)
u64 = ctypes.c_uint64


@pytest.fixture(scope="module")
def scene():
    o, d = synthetic_rays(1536, 0)
    _, bits = random_bitfield(0)
    bits = bits.numpy()
    N = len(o)
    nears, fars = np.empty(N, np.float32), np.empty(N, np.float32)
    kemu.call("tnl_near_far_from_aabb", o, d, AABB, N, 0.2, nears, fars, None)
    return o, d, bits, nears, fars


def _march_train(o, d, bits, nears, fars, noises, M, dt_gamma=0.0, max_steps=1024, fast=None):
    """fast = None: run BOTH forms of the second pass -- re-traversal (small workspace) and emission from the recorded sample
    parameters (workspace_fast) -- into buffers full of NaN, require bit-identical results, return them"""
    N = len(o)
    outs = []
    for f in ((False, True) if fast is None else (fast,)):
        xyzs, dirs, deltas = (np.full((M, w), np.nan, np.float32) for w in (3, 3, 2))      # the call zero-fills what no ray owns
        rays, counter = np.empty((N, 3), np.int32), np.zeros(2, np.int32)
        wsz = kemu.lib().tnl_march_rays_train_workspace_fast(N, max_steps) if f else kemu.lib().tnl_march_rays_train_workspace(N)
        ws = np.zeros(wsz, np.uint8)
        kemu.call("tnl_march_rays_train", o, d, bits, BOUND, dt_gamma, max_steps, N, CAS, H, M, nears, fars, xyzs, dirs, deltas,
                  rays, counter, noises, ws, wsz, None)
        outs.append((xyzs, dirs, deltas, rays, counter))
    if len(outs) == 2 and M > 0:
        for a, b in zip(*outs):
            assert np.array_equal(a.view(np.uint32) if a.dtype == np.float32 else a, b.view(np.uint32) if b.dtype == np.float32 else b)
    return outs[-1]


def _bits_equal(a, b):
    return np.array_equal(a.view(np.uint32), b.view(np.uint32))


# ------------------------------------------------------------------------------------------------ ray utilities
def test_near_far_morton_packbits(scene):
    o, d, bits, nears, fars = scene
    n_o, f_o = orc.near_far_from_aabb(o, d, AABB, 0.2)
    assert _bits_equal(nears, n_o) and _bits_equal(fars, f_o)
    rng = np.random.default_rng(1)
    c = rng.integers(0, 128, (20000, 3)).astype(np.int32)
    idx = np.empty(len(c), np.int32)
    kemu.call("tnl_morton3d", c, len(c), idx, None)
    assert np.array_equal(idx, orc.morton3D(c))
    back = np.empty_like(c)
    kemu.call("tnl_morton3d_invert", idx, len(c), back, None)
    assert np.array_equal(back, c)
    grid = rng.random(2 * 32 ** 3).astype(np.float32)
    for thresh in (0.0, 0.3, 0.999):
        pb = np.empty(grid.size // 8, np.uint8)
        kemu.call("tnl_packbits", grid, pb.size, thresh, pb, None)
        assert np.array_equal(pb, np.packbits(grid > thresh, bitorder='little'))
    sph = np.empty((len(o), 2), np.float32)
    kemu.call("tnl_sph_from_ray", o, d, 2.0, len(o), sph, None)
    assert np.abs(sph - orc.sph_from_ray(o, d, 2.0)).max() <= 4e-6       # atan2f / acosf of differently contracted arguments


@pytest.mark.parametrize("max_steps,dt_gamma", [(1024, 0.0), (256, 1.0 / 128)])
def test_march_rays_train_bitexact(scene, max_steps, dt_gamma):
    o, d, bits, nears, fars = scene
    N = len(o)
    noises = np.random.default_rng(3).random(N).astype(np.float32)
    M = N * 200
    xyzs, dirs, deltas, rays, counter = _march_train(o, d, bits, nears, fars, noises, M, dt_gamma, max_steps)
    x_o, d_o, l_o, r_o, c_o = orc.march_rays_train(o, d, BOUND, bits, CAS, H, nears, fars, noises, M, dt_gamma, max_steps)
    assert np.array_equal(counter, c_o) and int(c_o[0]) > N
    assert np.array_equal(rays, r_o)                       # ray order, offsets and counts (device-wide scan)
    assert _bits_equal(xyzs, x_o) and _bits_equal(dirs, d_o) and _bits_equal(deltas, l_o)


def test_march_train_empty_full_overflow(scene):
    o, d, bits, nears, fars = scene
    N = len(o)
    z = np.zeros(N, np.float32)
    *_, rays, counter = _march_train(o, d, np.zeros_like(bits), nears, fars, z, 128)
    assert counter[0] == 0 and counter[1] == N and rays[:, 2].sum() == 0
    full = np.full_like(bits, 255)
    x, _, _, rays, counter = _march_train(o, d, full, nears, fars, z, N * 64, 0.0, 64)
    r_o = orc.march_rays_train(o, d, BOUND, full, CAS, H, nears, fars, z, 0, 0.0, 64)[3]
    assert np.array_equal(rays, r_o) and rays[:, 2].max() == 64
    # overflow: M smaller than needed -> late rays dropped silently, earlier ones intact
    M = 128 * 51
    x2, _, _, rays2, _ = _march_train(o, d, full, nears, fars, z, M, 0.0, 64)
    assert np.array_equal(rays2, rays)
    keep = rays2[:, 1] + rays2[:, 2] <= M
    last = int((rays2[:, 1] + rays2[:, 2])[keep].max())
    assert _bits_equal(x2[:last], x[:last]) and np.abs(x2[last:]).sum() == 0.0


def test_composite_train_forward_backward(scene):
    o, d, bits, nears, fars = scene
    N = len(o)
    rng = np.random.default_rng(0)
    noises = rng.random(N).astype(np.float32)
    cnt = _march_train(o, d, bits, nears, fars, noises, 0)[4]
    M = int(cnt[0])
    xyzs, dirs, deltas, rays, _ = _march_train(o, d, bits, nears, fars, noises, M)
    sig = (rng.random(M) * 40).astype(np.float32)
    rgb = rng.random((M, 3)).astype(np.float32)
    ws, dp, im = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, 3), np.float32)
    kemu.call("tnl_composite_rays_train_forward", sig, rgb, deltas, rays, M, N, 1e-4, ws, dp, im, None)
    ws_o, dp_o, im_o = orc.composite_rays_train_forward(sig, rgb, deltas, rays, 1e-4)
    # warp-scan summation order differs from the oracle's sequential loop: fp32 reassociation over <= 1024 terms
    assert np.abs(ws - ws_o).max() <= 2e-5 and np.abs(im - im_o).max() <= 2e-5 and np.abs(dp - dp_o).max() <= 2e-4
    gws, gim = rng.standard_normal(N).astype(np.float32), rng.standard_normal((N, 3)).astype(np.float32)
    gs, gc = np.zeros(M, np.float32), np.zeros((M, 3), np.float32)
    kemu.call("tnl_composite_rays_train_backward", gws, gim, sig, rgb, deltas, rays, ws, im, M, N, 1e-4, gs, gc, None)
    gs_o, gc_o = orc.composite_rays_train_backward(gws, gim, sig, rgb, deltas, rays, ws_o, im_o, 1e-4)
    assert np.abs(gc - gc_o).max() <= 2e-5
    assert np.abs(gs - gs_o).max() / np.abs(gs_o).max() <= 1e-4


def test_inference_loop_march_composite_compact(scene):
    """renderer.py:342-368 driven with the emulated march_rays / composite_rays / compact_alive vs the C oracle"""
    o, d, bits, nears, fars = scene
    N = 700
    o, d, nears, fars = o[:N].copy(), d[:N].copy(), nears[:N].copy(), fars[:N].copy()
    ws, dp, im = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, 3), np.float32)
    ws_o, dp_o, im_o = ws.copy(), dp.copy(), im.copy()
    alive, rt = np.arange(N, dtype=np.int32), nears.copy()
    alive_o, rt_o = alive.copy(), rt.copy()
    rng = np.random.default_rng(0)
    n_alive, step, iters = N, 0, 0
    wsz = kemu.lib().tnl_compact_alive_workspace(N)
    work = np.zeros(max(wsz, 16), np.uint8)
    while step < 1024 and n_alive > 0:
        n_step = max(min(N // n_alive, 8), 1)
        M = n_alive * n_step
        M += 128 - M % 128
        x, dd, dl = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32)
        zeros = np.zeros(n_alive, np.float32)
        kemu.call("tnl_march_rays", n_alive, n_step, alive, rt, o, d, BOUND, 0.0, 1024, CAS, H, bits, nears, fars, x, dd, dl,
                  zeros, None)
        x_o, d_o, l_o = orc.march_rays(n_alive, n_step, alive_o, rt_o, o, d, BOUND, bits, CAS, H, nears, fars, zeros, 128, 0.0, 1024)
        assert _bits_equal(x, x_o) and _bits_equal(dl, l_o) and _bits_equal(dd, d_o)
        sig = (rng.random(M) * 30).astype(np.float32)
        rgb = rng.random((M, 3)).astype(np.float32)
        kemu.call("tnl_composite_rays", n_alive, n_step, 1e-4, alive, rt, sig, rgb, dl, ws, dp, im, None)
        orc.composite_rays(n_alive, n_step, alive_o, rt_o, sig, rgb, l_o, ws_o, dp_o, im_o, 1e-4)
        out, cnt = np.full(n_alive, -7, np.int32), np.zeros(1, np.int32)
        kemu.call("tnl_compact_alive", alive, n_alive, out, cnt, work, work.size, None)
        keep = alive[alive >= 0]
        assert int(cnt[0]) == len(keep) and np.array_equal(out[:len(keep)], keep)      # order kept
        alive = out[:len(keep)].copy()
        alive_o = alive_o[alive_o >= 0].copy()
        assert np.array_equal(alive, alive_o)
        n_alive = len(alive)
        step += n_step
        iters += 1
    assert iters > 10
    assert np.abs(ws - ws_o).max() <= 1e-5 and np.abs(im - im_o).max() <= 1e-5 and np.abs(dp - dp_o).max() <= 1e-4
    # n = 0 compaction: the count is written, nothing else touched
    cnt = np.full(1, 5, np.int32)
    kemu.call("tnl_compact_alive", None, 0, None, cnt, None, 0, None)
    assert cnt[0] == 0


def test_sh_encoder():
    g = torch.Generator().manual_seed(0)
    d = torch.randn(3000, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    out = np.empty((3000, 16), np.float32)
    kemu.call("tnl_sh_encode_forward", d.numpy(), out, 3000, 4, None)
    assert np.abs(out - of.sh16(d).numpy()).max() <= 1e-6
    assert abs(out[0, 0] - 0.28209479) < 1e-7


# ------------------------------------------------------------------------------------------------ plane sampling
def _cl(planes):
    """logical [3,C,R,R] -> stored channels-last [3][R][R][C]"""
    return np.ascontiguousarray(planes.permute(0, 2, 3, 1).numpy())


@pytest.mark.parametrize("C,R,fp16,use_perm", [(16, 64, False, False), (32, 96, False, True), (32, 96, True, True), (48, 40, True, False),
                                               (8, 32, False, True)])
def test_sampling_forward_backward(C, R, fp16, use_perm):
    g = torch.Generator().manual_seed(C + R)
    M = 3001
    planes = torch.randn(3, C, R, R, generator=g)
    xyz = (torch.rand(M, 3, generator=g) * 2 - 1) * BOUND
    xyz[:6] = torch.tensor([[-1.5, -1.5, -1.5], [1.5, 1.5, 1.5], [0, 0, 0], [1.5, -1.5, 0.3], [-1.4999, 1.4999, 0], [1.5, 0, 0]])
    inv_bound = float(np.float32(1.0) / np.float32(BOUND))
    nv = np.array([M - 100], np.int32)
    perm = None
    if use_perm:
        perm = np.empty(M, np.int32)
        wsz = kemu.lib().tnl_cell_sort_workspace(M, 16)
        work = np.zeros(max(wsz, 16), np.uint8)
        kemu.call("tnl_cell_sort", xyz.numpy(), M, nv, inv_bound, 16, perm, work, work.size, None)
        assert np.array_equal(np.sort(perm), np.arange(M))                       # a permutation
        assert np.array_equal(np.sort(perm[M - 100:]), np.arange(M - 100, M))     # rows >= n_valid last
    ref = of.sample_planes(planes, xyz, BOUND, fp16=fp16)
    feat = np.full((M, 3 * C), np.nan, np.float32)
    kemu.call("tnl_sample_planes_forward", _cl(planes), xyz.numpy(), M, R, C, inv_bound, int(fp16), nv, perm, feat, 0, None)
    assert np.abs(feat[:M - 100] - ref[:M - 100].numpy()).max() <= 1e-5 * ref.abs().max().item()
    assert np.abs(feat[M - 100:]).max() == 0.0                                    # skipped rows are zeros
    # fp16 feature stream: the same values rounded once
    feat_h = np.zeros((M, 3 * C), np.float16)
    kemu.call("tnl_sample_planes_forward", _cl(planes), xyz.numpy(), M, R, C, inv_bound, int(fp16), nv, perm, feat_h, 1, None)
    assert np.array_equal(feat_h[:M - 100], feat[:M - 100].astype(np.float16))
    # adjoint: <sample(P), G> == <P, scatter(G)> and scatter(G) == autograd of the oracle
    G = torch.randn(M, 3 * C, generator=g)
    G[M - 100:] = 0
    pl = planes.clone().requires_grad_(True)
    (of.sample_planes(pl, xyz, BOUND, fp16=fp16) * G).sum().backward()
    gp = np.zeros((3, R, R, C), np.float32)
    kemu.call("tnl_sample_planes_backward", G.numpy(), 0, xyz.numpy(), M, R, C, inv_bound, int(fp16), nv, perm, gp, None)
    want = _cl(pl.grad)
    assert np.abs(gp - want).max() <= 1e-4 * np.abs(want).max()
    gp_h = np.zeros((3, R, R, C), np.float32)
    Gh = G.numpy().astype(np.float16)
    kemu.call("tnl_sample_planes_backward", Gh, 1, xyz.numpy(), M, R, C, inv_bound, int(fp16), nv, perm, gp_h, None)
    pl.grad = None
    (of.sample_planes(pl, xyz, BOUND, fp16=fp16) * torch.from_numpy(Gh.astype(np.float32))).sum().backward()
    assert np.abs(gp_h - _cl(pl.grad)).max() <= 1e-4 * np.abs(want).max()
    # plane by plane (the multi-GPU step overlaps the exchange of one plane with the scatter of the next): same per-plane sums,
    # and plane p's call touches plane p only
    if C in (16, 32, 48):
        gp_p = np.zeros((3, R, R, C), np.float32)
        for p in range(3):
            before = gp_p.copy()
            kemu.call("tnl_sample_planes_backward_plane", Gh, 1, xyz.numpy(), M, R, C, inv_bound, int(fp16), nv, perm, gp_p, p, None)
            others = [q for q in range(3) if q != p]
            assert np.array_equal(gp_p[others], before[others])
        assert np.abs(gp_p - gp_h).max() <= 1e-5 * np.abs(want).max()
        assert kemu.lib().tnl_sample_planes_backward_plane(kemu.p(Gh), 1, kemu.p(xyz.numpy()), M, R, C, ctypes.c_float(inv_bound), int(fp16),
                                                           None, None, kemu.p(gp_p), 3, None) == -1


def test_cell_sort_orders_by_cell():
    g = torch.Generator().manual_seed(5)
    M, G = 5000, 8
    xyz = ((torch.rand(M, 3, generator=g) * 2 - 1) * BOUND).numpy()
    perm = np.empty(M, np.int32)
    wsz = kemu.lib().tnl_cell_sort_workspace(M, G)
    work = np.zeros(max(wsz, 16), np.uint8)
    inv_bound = float(np.float32(1.0) / np.float32(BOUND))
    kemu.call("tnl_cell_sort", xyz, M, None, inv_bound, G, perm, work, work.size, None)
    assert np.array_equal(np.sort(perm), np.arange(M))
    cell = np.clip(np.floor((xyz[perm] * np.float32(inv_bound) + 1) * 0.5 * G), 0, G - 1).astype(np.int64)
    # points of one cell are contiguous in the visit order: the number of cell changes equals the number of non-empty cells - 1
    key = (cell[:, 0] * G + cell[:, 1]) * G + cell[:, 2]
    changes = int((key[1:] != key[:-1]).sum())
    assert changes == len(np.unique(key)) - 1


# ------------------------------------------------------------------------------------------------ dirty tiles
def test_dirty_tiles_cover_every_sample_and_pack_roundtrip(scene):
    """the property the work-list IDWT and the sparse gradient exchange rely on: every texel a sample of an occupied cell
    reads or writes lies in a marked tile -- checked on the samples the (emulated) marcher actually produces, with and
    without the fp16 rounding of the projected coordinates."""
    from trinerflet_b200 import scene as sc
    o, d, _, nears, fars = scene
    bits = sc.packbits_cpu(sc.ball_density_grid(BOUND, 0.45, 1.0, H), 0.5).numpy()       # a clean ball: most tiles stay unmarked
    N = len(o)
    noises = np.random.default_rng(7).random(N).astype(np.float32)
    cnt = _march_train(o, d, bits, nears, fars, noises, 0)[4]
    M = int(cnt[0])
    xyzs = _march_train(o, d, bits, nears, fars, noises, M)[0]
    R, T, C = 512, 32, 8
    nt = R // T
    flags = np.zeros((3, nt, nt), np.uint8)
    kemu.call("tnl_mark_dirty_tiles", bits, CAS, H, BOUND, R, T, 2, flags, None)
    assert 0.02 < flags.mean() < 0.5 and M > 10000
    for fp16 in (False, True):
        u = of.project_coords(torch.from_numpy(xyzs), BOUND, fp16=fp16).numpy().astype(np.float64)
        ix = np.clip((u + 1) * 0.5 * (R - 1), 0, R - 1)
        lo, hi = np.floor(ix).astype(np.int64), np.minimum(np.floor(ix).astype(np.int64) + 1, R - 1)
        for p, (a, b) in enumerate(of.PLANE_AXES):
            for tx in (lo[:, a] // T, hi[:, a] // T):
                for ty in (lo[:, b] // T, hi[:, b] // T):
                    assert flags[p, ty, tx].all()
    # pack -> unpack (fp32 transport) restores exactly the marked tiles, scaled; bf16 transport rounds to 8 bits
    ids = np.flatnonzero(flags.reshape(-1)).astype(np.int32)
    rng = np.random.default_rng(0)
    planes = rng.standard_normal((3, R, R, C)).astype(np.float32)
    compact = np.zeros((len(ids), T, T, C), np.float32)
    kemu.call("tnl_tiles_pack", planes, ids, len(ids), R, C, T, compact, 0, None)
    t0 = int(ids[0])
    p0, ty0, tx0 = t0 // (nt * nt), (t0 // nt) % nt, t0 % nt
    assert np.array_equal(compact[0], planes[p0, ty0 * T:(ty0 + 1) * T, tx0 * T:(tx0 + 1) * T])
    back = np.zeros_like(planes)
    kemu.call("tnl_tiles_unpack", compact, ids, len(ids), R, C, T, 0.5, 0, back, None)
    mask = np.repeat(np.repeat(flags.astype(bool), T, axis=1), T, axis=2)[..., None]
    assert np.array_equal(back, np.where(mask, planes * 0.5, 0).astype(np.float32))
    compact16 = np.zeros((len(ids), T, T, C), np.uint16)
    kemu.call("tnl_tiles_pack", planes, ids, len(ids), R, C, T, compact16, 1, None)
    back16 = np.zeros_like(planes)
    kemu.call("tnl_tiles_unpack", compact16, ids, len(ids), R, C, T, 1.0, 1, back16, None)
    want16 = torch.from_numpy(planes).bfloat16().float().numpy()
    assert np.array_equal(back16, np.where(mask, want16, 0).astype(np.float32))
    # tile-wise zero fill touches only the listed tiles
    buf = np.ones_like(planes)
    count = np.array([len(ids) - 3], np.int32)
    kemu.call("tnl_tiles_zero", buf, ids, count, len(ids), R, C, T, None)
    flags2 = flags.reshape(-1).copy()
    flags2[ids[-3:]] = 0
    mask2 = np.repeat(np.repeat(flags2.reshape(3, nt, nt).astype(bool), T, axis=1), T, axis=2)[..., None]
    assert np.array_equal(buf, np.where(mask2, 0, 1).astype(np.float32) * np.ones_like(planes))


# ------------------------------------------------------------------------------------------------ optimizer epilogue
def test_fused_adam_kernels_match_torch_adam():
    g = torch.Generator().manual_seed(0)
    n = 10007
    p0 = torch.randn(n, generator=g)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    p, m, v = p0.numpy().copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    state = np.zeros(4, np.float32)                     # {step, 1 - beta1^step, sqrt(1 - beta2^step)}: starts at step 0 (optim.py)
    scale = 1024.0
    inv_scale = np.array([1.0 / scale], np.float32)
    for it in range(4):
        grad = torch.randn(n, generator=g)
        found = np.zeros(1, np.float32)
        gs = (grad * scale).numpy().copy()
        if it == 2:
            gs[123] = np.inf                            # overflow step: skipped, state untouched
        kemu.call("tnl_grad_nonfinite", gs, u64(n), found, None)
        assert found[0] == (1.0 if it == 2 else 0.0)
        before = (p.copy(), m.copy(), v.copy(), state.copy())
        kemu.call("tnl_adam_prepare", state, found, 0.9, 0.99, None)
        kemu.call("tnl_adam_step", p, gs, m, v, u64(n), inv_scale, found, state, 1e-2, 0.9, 0.99, 1e-15, 0.0, None)
        if it == 2:
            assert np.array_equal(p, before[0]) and np.array_equal(m, before[1]) and np.array_equal(v, before[2])
            assert state[0] == before[3][0]
            continue
        ref.grad = grad.clone()
        opt.step()
        assert np.abs(p - ref.detach().numpy()).max() <= 2e-6
    assert state[0] == 3.0                              # three applied steps, one skipped


def test_grid_cell_positions_and_ema():
    rng = np.random.default_rng(0)
    n = 4096
    coords = rng.integers(0, 128, (n, 3)).astype(np.int32)
    idx = orc.morton3D(coords)
    noise = rng.random((n, 3)).astype(np.float32)
    xyz = np.empty((n, 3), np.float32)
    bound_c = 1.5
    kemu.call("tnl_grid_cell_positions", idx, n, 128, bound_c, noise, xyz, None)
    # renderer.py:474-483 with torch's CUDA scalar rules, fp32
    t = torch.from_numpy(coords).float()
    hgs = bound_c / 128
    want = (2 * t * torch.tensor(1.0 / 127, dtype=torch.float32) - 1) * (bound_c - hgs) + (torch.from_numpy(noise) * 2 - 1) * hgs
    assert np.abs(xyz - want.numpy()).max() <= 2e-7
    grid = rng.random(n).astype(np.float32)
    grid[::7] = -1.0                                    # untrained cells stay untouched
    tmp = rng.random(n).astype(np.float32)
    tmp[::5] = -1.0
    want = np.where((grid >= 0) & (tmp >= 0), np.maximum(grid * np.float32(0.95), tmp), grid)
    acc = np.full(1, 123.0, np.float64)                 # (zeroed by the call)
    kemu.call("tnl_grid_ema_update_sum", grid, tmp, n, 0.95, acc, None)
    assert np.array_equal(grid, want)
    assert abs(acc[0] - np.clip(want, 0, None).astype(np.float64).sum()) <= 1e-3
    # mean / threshold / packbits on the device (renderer.py:528-534)
    for cap in (10.0, 0.2):
        mean, bits = np.zeros(1, np.float32), np.zeros(n // 8, np.uint8)
        kemu.call("tnl_packbits_mean", grid, n, acc, cap, mean, bits, None)
        assert abs(mean[0] - np.clip(want, 0, None).mean()) <= 1e-6
        assert np.array_equal(bits, np.packbits(want > min(mean[0], np.float32(cap)), bitorder="little"))
    # tmp_grid[indices] = sigma * density_scale
    sel = rng.permutation(n)[:1000].astype(np.int32)
    sig = rng.random(1000).astype(np.float32)
    tmp2 = np.full(n, -1.0, np.float32)
    kemu.call("tnl_grid_scatter", sel, sig, 1000, 2.0, tmp2, None)
    want2 = np.full(n, -1.0, np.float32)
    want2[sel] = sig * np.float32(2.0)
    assert np.array_equal(tmp2, want2)


# ------------------------------------------------------------------------------------------------ step feeder
def test_rays_from_ids_kernel_matches_reference_golden(golden_dir):
    gold = np.load(os.path.join(golden_dir, "rays_ref.npz"))
    H, W, bs = 37, 53, int(gold["C_bs"])
    poses = np.ascontiguousarray(gold["B_poses"], np.float32)
    intr = [float(v) for v in gold["B_intr"]]
    images = np.ascontiguousarray(gold["C_images"], np.float32)
    for b in (0, 3, 7):
        ids = np.ascontiguousarray(gold["C_perm"][b * bs:(b + 1) * bs], np.int64)
        n = len(ids)
        ro, rd, gt = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32), np.empty((n, 4), np.float32)
        kemu.call("tnl_rays_from_ids", poses, 4, *intr, H, W, ids, 0, n, images, 4, ro, rd, gt, None)
        assert np.array_equal(ro, gold[f"C_b{b}_rays_o"]) and np.array_equal(rd, gold[f"C_b{b}_rays_d"])
        assert np.array_equal(gt, gold[f"C_b{b}_images"])
    n = 4 * H * W
    ro, rd = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
    kemu.call("tnl_rays_from_ids", poses, 4, *intr, H, W, None, 0, n, None, 0, ro, rd, None, None)
    assert np.array_equal(rd.reshape(4, -1, 3), gold["B_full_d"]) and np.array_equal(ro.reshape(4, -1, 3), gold["B_full_o"])


# ------------------------------------------------------------------------------------------------ IDWT (dense + work-list)
def _cl_coefs(t):
    """logical [3,C,3,n,n] -> stored [3][3][n][n][C]"""
    return np.ascontiguousarray(t.permute(0, 2, 3, 4, 1).numpy())


def _idwt_chain_emu(pf, coefs, plan=None, parts=3):
    """the level loop of _BuildPlanes.forward on the emulated entry points -> (planes [3][R][R][C], abs_sums [L])"""
    x = _cl(pf)
    C, n = x.shape[3], x.shape[1]
    abs_sums = np.zeros(len(coefs), np.float32)
    for l, yh in enumerate(coefs):
        out = np.full((3, 2 * n, 2 * n, C), np.nan, np.float32)
        a = abs_sums[l:l + 1]
        if plan is None:
            kemu.call("tnl_idwt_level_forward", x, _cl_coefs(yh), out, n, C, a, None)
        else:
            s = plan.fwd[l]
            kemu.call("tnl_idwt_level_forward_sparse", x, _cl_coefs(yh), out, n, C, a, s["active"].numpy(), s["clean"].numpy(),
                      s["counts"].numpy(), s["cap_active"], s["cap_clean"], parts, None)
        x, n = out, 2 * n
    return x, abs_sums


def _idwt_chain_bwd_emu(gout, coefs, n0, reg, plan=None):
    """adjoint chain (top level first) with the fused regulariser gradient -> (g_x0, [g_yh_l])"""
    L = len(coefs)
    g = gout
    C = g.shape[3]
    g_yh = [None] * L
    reg_grad = np.array([reg], np.float32)
    for l in reversed(range(L)):
        n = n0 * 2 ** l
        gx = np.full((3, n, n, C), np.nan, np.float32)
        gy = np.full((3, 3, n, n, C), np.nan, np.float32)
        yh = _cl_coefs(coefs[l])
        if plan is None:
            kemu.call("tnl_idwt_level_backward", g, gx, gy, n, C, yh, reg_grad, 0.25, 0, 3, None)
        else:
            s = plan.bwd[l]
            kemu.call("tnl_idwt_level_backward_sparse", g, gx, gy, n, C, yh, reg_grad, 0.25, s["active"].numpy(), s["clean"].numpy(),
                      s["counts"].numpy(), s["cap_active"], s["cap_clean"], 3, None, None)
        g_yh[l] = gy
        g = gx
    return g, g_yh


@pytest.mark.parametrize("C,n0,levels", [(8, 8, 2), (16, 16, 1), (24, 24, 1)])
def test_idwt_dense_entry_points_match_oracle(C, n0, levels):
    from oracle import wavelet as ow
    g = torch.Generator().manual_seed(C + n0)
    pf = torch.randn(3, C, n0, n0, generator=g, requires_grad=True)
    coefs = [torch.randn(3, C, 3, n0 * 2 ** l, n0 * 2 ** l, generator=g, requires_grad=True) for l in range(levels)]
    ref = ow.build_planes(pf, coefs)
    planes, abs_sums = _idwt_chain_emu(pf.detach(), [c.detach() for c in coefs])
    assert np.abs(planes - _cl(ref.detach())).max() <= 1e-5 * ref.abs().max().item()
    for l, c in enumerate(coefs):
        assert abs(abs_sums[l] - c.detach().abs().sum().item()) <= 1e-4 * abs_sums[l]
    gout = torch.randn(ref.shape, generator=g)
    lam = 0.5
    (ref * gout).sum().backward(retain_graph=True)
    gx, gyh = _idwt_chain_bwd_emu(_cl(gout), [c.detach() for c in coefs], n0, lam)
    assert np.abs(gx - _cl(pf.grad)).max() <= 1e-4 * pf.grad.abs().max().item()
    for l, c in enumerate(coefs):
        want = c.grad + 0.25 * lam * torch.sign(c.detach())         # + reg_coef * (*reg_grad) * sign(yh)
        assert np.abs(gyh[l] - _cl_coefs(want)).max() <= 1e-4 * want.abs().max().item()


@pytest.mark.parametrize("C,n0,levels,density", [(16, 16, 2, 0.08), (8, 16, 1, 0.3), (16, 16, 1, 0.0)])
def test_worklist_idwt_entry_points_equal_dense(C, n0, levels, density):
    """the work-list entry points driven by idwt_plan.IdwtPlan (built on the CPU): bit-identical to the dense entry points
    inside the marked tiles; |yh| sums complete; the adjoint of a gradient that vanishes outside the marked tiles is
    bit-identical everywhere, regulariser gradient included (the CPU twin of tests/test_gpu_encoder.py)."""
    from trinerflet_b200.idwt_plan import IdwtPlan
    R = n0 * 2 ** levels
    T = R // 32
    g = torch.Generator().manual_seed(5)
    flags = torch.rand(3, T, T, generator=g) < density
    if density > 0:
        flags[:, T // 4: T // 2 + 1, T // 3: T // 2 + 1] = True
    plan = IdwtPlan(R, n0, levels, C, "cpu").update(flags)
    pf = torch.randn(3, C, n0, n0, generator=g)
    coefs = [0.1 * torch.randn(3, C, 3, n0 * 2 ** l, n0 * 2 ** l, generator=g) for l in range(levels)]
    coefs[-1][:, :, :, ::3, ::4] = 0.0                               # exact zeros: sign(0) = 0
    dense, abs_d = _idwt_chain_emu(pf, coefs)
    sparse, abs_s = _idwt_chain_emu(pf, coefs, plan)
    mask = np.repeat(np.repeat(flags.numpy(), 32, axis=1), 32, axis=2)[..., None]
    assert np.array_equal(np.where(mask, sparse, 0).view(np.uint32), np.where(mask, dense, 0).view(np.uint32))
    assert np.abs(abs_s - abs_d).max() <= 1e-5 * abs_d.max()
    gout = (np.random.default_rng(1).standard_normal(dense.shape).astype(np.float32) * mask).astype(np.float32)
    gx_d, gy_d = _idwt_chain_bwd_emu(gout, coefs, n0, 0.7)
    gx_s, gy_s = _idwt_chain_bwd_emu(gout, coefs, n0, 0.7, plan)
    assert np.array_equal(gx_s.view(np.uint32), gx_d.view(np.uint32))
    for a, b in zip(gy_s, gy_d):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


# ------------------------------------------------------------------------------------------------ device-driven inference loop
def _host_driven_loop(o, d, bits, nears, fars, field, max_steps, T_thresh=1e-4):
    """renderer.py:342-368 with the host-driven entry points (fresh zeroed buffers and a host read of the count per iteration)"""
    N = len(o)
    ws, dp, im = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, 3), np.float32)
    alive, rt = np.arange(N, dtype=np.int32), nears.copy()
    work = np.zeros(max(kemu.lib().tnl_compact_alive_workspace(N), 16), np.uint8)
    n_alive, step, iters = N, 0, 0
    while step < max_steps and n_alive > 0:
        n_step = max(min(N // n_alive, 8), 1)
        M = n_alive * n_step
        M += 128 - M % 128
        x, dd, dl = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32)
        kemu.call("tnl_march_rays", n_alive, n_step, alive, rt, o, d, BOUND, 0.0, max_steps, CAS, H, bits, nears, fars, x, dd, dl,
                  np.zeros(n_alive, np.float32), None)
        sig, rgb = field(x, dd)
        kemu.call("tnl_composite_rays", n_alive, n_step, T_thresh, alive, rt, sig, rgb, dl, ws, dp, im, None)
        out, cnt = np.zeros(n_alive, np.int32), np.zeros(1, np.int32)
        kemu.call("tnl_compact_alive", alive, n_alive, out, cnt, work, work.size, None)
        n_alive = int(cnt[0])
        alive = out[:n_alive].copy()
        step += n_step
        iters += 1
    return ws, dp, im, rt, iters


def _device_driven_loop(o, d, bits, nears, fars, field, max_steps, chunk, T_thresh=1e-4):
    """the same loop with the state in `ctrl`: `chunk` iterations are issued between two reads of the state"""
    N = len(o)
    ws, dp, im = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, 3), np.float32)
    lists = [np.arange(N, dtype=np.int32), np.full(N, -9, np.int32)]
    rt = nears.copy()
    rows = N + 128 - N % 128
    # persistent buffers, deliberately dirty: the marcher must clear every row it owns
    x, dd, dl = (np.full((rows, k), 7.0, np.float32) for k in (3, 3, 2))
    ctrl = np.zeros(8, np.int32)
    ctrl[6] = N
    work = np.zeros(max(kemu.lib().tnl_compact_alive_workspace(N), 16), np.uint8)
    cap, issued, reads = N, 0, 0
    while True:
        for _ in range(chunk):
            kemu.call("tnl_infer_plan", ctrl, N, max_steps, None)
            kemu.call("tnl_march_rays_dev", ctrl, cap, lists[0], rt, o, d, BOUND, 0.0, max_steps, CAS, H, bits, fars, x, dd, dl, None, None)
            n_rows = int(ctrl[3])                  # (the field kernels read this through their n_valid pointer)
            sig, rgb = field(x[:max(n_rows, 1)], dd[:max(n_rows, 1)])
            kemu.call("tnl_composite_rays_dev", ctrl, cap, T_thresh, lists[0], rt, sig, rgb, dl, ws, dp, im, None)
            kemu.call("tnl_compact_alive_dev", ctrl, cap, lists[0], lists[1], work, work.size, None)
            lists.reverse()
            issued += 1
        reads += 1                                  # the host looks at the state once per chunk
        assert ctrl[6] <= cap
        cap = int(ctrl[6])
        if cap == 0 or ctrl[2] >= max_steps:
            break
    return ws, dp, im, rt, int(ctrl[4]), issued, reads


@pytest.mark.parametrize("max_steps,chunk", [(1024, 8), (48, 5)])
def test_device_driven_inference_loop_equals_host_driven(scene, max_steps, chunk):
    o, d, bits, nears, fars = scene
    N = 500
    o, d, nears, fars = o[:N].copy(), d[:N].copy(), nears[:N].copy(), fars[:N].copy()

    def field(x, dd):      # a deterministic stand-in for the sigma / colour heads (a function of the sample only)
        s = (np.abs(np.sin(37.0 * x[:, 0] + 11.0 * x[:, 1])) * 12.0).astype(np.float32)
        c = np.abs(np.cos(x * 5.0 + dd)).astype(np.float32)
        return s, np.ascontiguousarray(c)

    ws_h, dp_h, im_h, rt_h, it_h = _host_driven_loop(o, d, bits, nears, fars, field, max_steps)
    ws_d, dp_d, im_d, rt_d, it_d, issued, reads = _device_driven_loop(o, d, bits, nears, fars, field, max_steps, chunk)
    assert it_d == it_h and issued >= it_h and reads == -(-issued // chunk) and reads < it_h
    assert _bits_equal(ws_d, ws_h) and _bits_equal(dp_d, dp_h) and _bits_equal(im_d, im_h) and _bits_equal(rt_d, rt_h)
    if max_steps == 48:
        assert it_h < 48     # the step budget, not the alive count, ended this one


# ------------------------------------------------------------------------------------------------ edge cases
def test_edge_cases_empty_single_and_degenerate_inputs(scene):
    """empty and one-element inputs, rays that miss the box, zero / out-of-range n_valid, dead and full alive lists, tails
    of the vectorised kernels -- the cases where index arithmetic usually breaks (run under tests/emu/run_asan.sh too)"""
    o, d, bits, nears, fars = scene
    z = None
    # N = 0 everywhere: nothing is launched, nothing is touched
    e3, e1 = np.zeros((0, 3), np.float32), np.zeros(0, np.float32)
    kemu.call("tnl_near_far_from_aabb", e3, e3, AABB, 0, 0.2, e1, e1, z)
    kemu.call("tnl_packbits", e1, 0, 0.5, np.zeros(0, np.uint8), z)
    kemu.call("tnl_sample_planes_forward", np.zeros((3, 4, 4, 8), np.float32), e3, 0, 4, 8, 1.0, 0, z, z, np.zeros((0, 24), np.float32), 0, z)
    kemu.call("tnl_tiles_pack", np.zeros((3, 32, 32, 8), np.float32), np.zeros(0, np.int32), 0, 32, 8, 32, np.zeros(0, np.float32), 0, z)
    # one ray; a ray that points away from the box (far < near -> no samples, count 0)
    o1 = np.array([[0.0, 0.0, 4.0], [0.0, 0.0, 4.0]], np.float32)
    d1 = np.array([[0.0, 0.0, -1.0], [0.0, 0.0, 1.0]], np.float32)
    n1, f1 = np.empty(2, np.float32), np.empty(2, np.float32)
    kemu.call("tnl_near_far_from_aabb", o1, d1, AABB, 2, 0.2, n1, f1, z)
    n_o, f_o = orc.near_far_from_aabb(o1, d1, AABB, 0.2)
    assert _bits_equal(n1, n_o) and _bits_equal(f1, f_o) and n1[0] == 2.5 and f1[0] == 5.5 and f1[1] < n1[1]   # box behind the ray
    full = np.full_like(bits, 255)
    for n in (1, 2):
        x, _, dl, rays, cnt = _march_train(o1[:n], d1[:n], full, n1[:n], f1[:n], np.zeros(n, np.float32), 2048)
        r_o = orc.march_rays_train(o1[:n], d1[:n], BOUND, full, CAS, H, n1[:n], f1[:n], np.zeros(n, np.float32), 2048)[3]
        assert np.array_equal(rays, r_o) and rays[0, 2] > 100 and cnt[1] == n
        if n == 2:
            assert rays[1, 2] == 0
    # max_steps = 1: exactly one sample per hitting ray
    *_, rays, cnt = _march_train(o1, d1, full, n1, f1, np.zeros(2, np.float32), 128, 0.0, 1)
    assert rays[0, 2] == 1 and rays[1, 2] == 0 and cnt[0] == 1
    # compositing: a ray without samples, zero density, and a wall that saturates at the first sample
    rays3 = np.array([[0, 0, 0], [1, 0, 4], [2, 4, 4]], np.int32)
    deltas = np.full((8, 2), 0.01, np.float32)
    sig = np.concatenate([np.zeros(4), np.full(4, 1e6)]).astype(np.float32)
    rgb = np.full((8, 3), 0.5, np.float32)
    ws, dp, im = np.full(3, -1, np.float32), np.full(3, -1, np.float32), np.full((3, 3), -1, np.float32)
    kemu.call("tnl_composite_rays_train_forward", sig, rgb, deltas, rays3, 8, 3, 1e-4, ws, dp, im, z)
    ws_o, dp_o, im_o = orc.composite_rays_train_forward(sig, rgb, deltas, rays3, 1e-4)
    assert np.allclose(ws, ws_o, atol=1e-6) and np.allclose(im, im_o, atol=1e-6) and np.allclose(dp, dp_o, atol=1e-6)
    assert ws[0] == 0 and ws[1] == 0 and abs(ws[2] - 1) < 1e-6
    gs, gc = np.full(8, 9, np.float32), np.full((8, 3), 9, np.float32)
    kemu.call("tnl_composite_rays_train_backward", np.ones(3, np.float32), np.ones((3, 3), np.float32), sig, rgb, deltas, rays3, ws, im,
              8, 3, 1e-4, gs, gc, z)
    gs_o, gc_o = orc.composite_rays_train_backward(np.ones(3, np.float32), np.ones((3, 3), np.float32), sig, rgb, deltas, rays3, ws_o,
                                                   im_o, 1e-4)
    assert np.allclose(gs[:5], gs_o[:5], atol=1e-5) and np.allclose(gc[:5], gc_o[:5], atol=1e-6)
    # alive-list compaction: all dead, all alive, a single element
    work = np.zeros(64, np.uint8)
    for alive in (np.full(37, -1, np.int32), np.arange(37, dtype=np.int32), np.array([5], np.int32), np.array([-1], np.int32)):
        out, cnt = np.full(len(alive), -7, np.int32), np.full(1, -7, np.int32)
        kemu.call("tnl_compact_alive", alive, len(alive), out, cnt, work, work.size, z)
        keep = alive[alive >= 0]
        assert cnt[0] == len(keep) and np.array_equal(out[:len(keep)], keep) and (out[len(keep):] == -7).all()
    # sampling: n_valid = 0, n_valid beyond M, negative n_valid; points outside the box clamp to the border texels
    g = torch.Generator().manual_seed(3)
    planes = torch.randn(3, 8, 16, 16, generator=g)
    xyz = torch.tensor([[9.0, -9.0, 0.3], [-1.5, 1.5, 1.5], [0.1, 0.2, 0.3]])
    ref = of.sample_planes(planes, xyz, BOUND).numpy()
    inv_bound = float(np.float32(1.0) / np.float32(BOUND))
    for nv, rows in ((0, 0), (-4, 0), (2, 2), (99, 3)):
        feat = np.full((3, 24), np.nan, np.float32)
        kemu.call("tnl_sample_planes_forward", _cl(planes), xyz.numpy(), 3, 16, 8, inv_bound, 0, np.array([nv], np.int32), z, feat, 0, z)
        assert np.allclose(feat[:rows], ref[:rows], atol=1e-6) and (feat[rows:] == 0).all()
        gp = np.zeros((3, 16, 16, 8), np.float32)
        kemu.call("tnl_sample_planes_backward", np.ones((3, 24), np.float32), 0, xyz.numpy(), 3, 16, 8, inv_bound, 0,
                  np.array([nv], np.int32), z, gp, z)
        assert abs(gp.sum() - rows * 24) <= 1e-4                       # bilinear weights of a point sum to one per channel
    # cell sort: one point; many points in one cell
    for pts in (np.zeros((1, 3), np.float32), np.full((300, 3), 0.01, np.float32)):
        perm = np.full(len(pts), -1, np.int32)
        wsz = kemu.lib().tnl_cell_sort_workspace(len(pts), 64)
        wk = np.zeros(max(wsz, 16), np.uint8)
        kemu.call("tnl_cell_sort", pts, len(pts), z, inv_bound, 64, perm, wk, wk.size, z)
        assert np.array_equal(np.sort(perm), np.arange(len(pts)))
    # optimizer: lengths below / off the 4-wide vector path
    for n in (1, 3, 5, 1023):
        p, m, v = np.ones(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
        gr = np.full(n, 0.5, np.float32)
        state = np.zeros(4, np.float32)
        kemu.call("tnl_adam_prepare", state, z, 0.9, 0.99, z)
        kemu.call("tnl_adam_step", p, gr, m, v, u64(n), z, z, state, 1e-2, 0.9, 0.99, 1e-15, 0.0, z)
        assert np.allclose(p, 1 - 1e-2, atol=1e-6) and np.allclose(m, 0.05, atol=1e-7)   # first Adam step moves by lr * sign(g)
        found = np.zeros(1, np.float32)
        gr[n - 1] = np.nan
        kemu.call("tnl_grad_nonfinite", gr, u64(n), found, z)
        assert found[0] == 1.0
    # IDWT: smallest supported level; unsupported geometry is an argument error, not a launch
    lib = kemu.lib()
    x, yh, out = np.zeros((3, 8, 8, 8), np.float32), np.zeros((3, 3, 8, 8, 8), np.float32), np.full((3, 16, 16, 8), np.nan, np.float32)
    x[:] = 1.0
    kemu.call("tnl_idwt_level_forward", x, yh, out, 8, 8, z, z)
    assert abs(out[:, 8, 8, :].mean() - 1.0) < 1e-5                    # DC gain 1 in the interior (IDWT(2*c, 0) = c)
    assert lib.tnl_idwt_level_forward(kemu.p(x), kemu.p(yh), kemu.p(out), 12, 8, None, None) == -1
    assert lib.tnl_idwt_level_forward(kemu.p(x), kemu.p(yh), kemu.p(out), 8, 12, None, None) == -1


# ------------------------------------------------------------------------------------------------ MLP heads (mma.sync kernels)
def _mlp_pack_emu(C, hidden, W):
    from trinerflet_b200._lib import MlpDims
    dims = MlpDims(3 * C, hidden, hidden)
    nbytes = kemu.lib().tnl_mlp_packed_bytes(ctypes.byref(dims))
    assert nbytes > 0
    packed = np.zeros(nbytes, np.uint8)
    kemu.call("tnl_mlp_pack_weights", ctypes.byref(dims), *[kemu.f32(w.numpy()) for w in W], packed, None)
    return dims, packed


@pytest.mark.parametrize("C,hidden,M", [(16, 64, 333), (32, 64, 130), (48, 128, 70)])
def test_mlp_forward_mma_kernels_match_fp16_oracle(C, hidden, M):
    """csrc/mlp.cu (mma.sync.m16n8k16 + fragment-ordered weights; the host build emulates the warp-wide instructions)
    against the oracle's fp16-autocast emulation: same tolerances as tests/test_gpu_field.py"""
    g = torch.Generator().manual_seed(C + hidden)
    W = of.init_mlp_weights(C, hidden, hidden, gen=g)
    feat = 0.5 * torch.randn(M, 3 * C, generator=g)
    d = torch.randn(M, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    s_o, rgb_o, geo_o = of.mlp_forward(feat, d, W, fp16=True)
    dims, packed = _mlp_pack_emu(C, hidden, W)
    sigma, rgb, geo = np.full(M, np.nan, np.float32), np.full((M, 3), np.nan, np.float32), np.full((M, 15), np.nan, np.float32)
    kemu.call("tnl_mlp_forward", ctypes.byref(dims), packed, feat.numpy(), 0, d.numpy(), M, None, sigma, rgb, geo, None)
    assert np.abs(rgb - rgb_o.numpy()).max() <= 2e-3
    assert (np.abs(sigma - s_o.numpy()) / np.maximum(np.abs(s_o.numpy()), 1e-3)).max() <= 4e-3
    assert np.abs(geo - geo_o.numpy()).max() <= 2e-3 * max(1.0, geo_o.abs().max().item())
    # fp16 feature stream: the same rounding point, identical outputs; n_valid: skipped rows are zeros
    sigma_h, rgb_h = np.zeros(M, np.float32), np.zeros((M, 3), np.float32)
    kemu.call("tnl_mlp_forward", ctypes.byref(dims), packed, feat.numpy().astype(np.float16), 1, d.numpy(), M, None, sigma_h, rgb_h,
              None, None)
    assert np.array_equal(sigma_h, sigma) and np.array_equal(rgb_h, rgb)
    nv = np.array([M // 3], np.int32)
    sigma_v, rgb_v = np.full(M, np.nan, np.float32), np.full((M, 3), np.nan, np.float32)
    kemu.call("tnl_mlp_forward", ctypes.byref(dims), packed, feat.numpy(), 0, d.numpy(), M, nv, sigma_v, rgb_v, None, None)
    assert np.array_equal(sigma_v[:M // 3], sigma[:M // 3]) and (sigma_v[M // 3:] == 0).all() and (rgb_v[M // 3:] == 0).all()
    # density only (dirs = NULL)
    sigma_d, geo_d = np.zeros(M, np.float32), np.zeros((M, 15), np.float32)
    kemu.call("tnl_mlp_forward", ctypes.byref(dims), packed, feat.numpy(), 0, None, M, None, sigma_d, None, geo_d, None)
    assert np.array_equal(sigma_d, sigma) and np.array_equal(geo_d, geo)


@pytest.mark.parametrize("C,M", [(16, 300), (32, 150), (48, 77)])
def test_mlp_backward_mma_kernels_match_fp16_oracle(C, M):
    g = torch.Generator().manual_seed(3 + C)
    W = of.init_mlp_weights(C, 64, 64, gen=g)
    feat = 0.5 * torch.randn(M, 3 * C, generator=g)
    d = torch.randn(M, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    gs = torch.randn(M, generator=g) * 64.0            # loss-scaled gradients, as GradScaler produces
    grgb = torch.randn(M, 3, generator=g) * 64.0
    W_o = [w.clone().requires_grad_(True) for w in W]
    f_o = feat.clone().requires_grad_(True)
    s_o, rgb_o, _ = of.mlp_forward(f_o, d, W_o, fp16=True)
    ((s_o * gs).sum() + (rgb_o * grgb).sum()).backward()
    dims, packed = _mlp_pack_emu(C, 64, W)
    g_feat = np.full((M, 3 * C), np.nan, np.float32)
    gW = [np.zeros(tuple(w.shape), np.float32) for w in W]
    kemu.call("tnl_mlp_backward", ctypes.byref(dims), packed, feat.numpy(), 0, d.numpy(), M, None, gs.numpy(), grgb.numpy(), g_feat,
              *gW, None)
    assert np.linalg.norm(g_feat - f_o.grad.numpy()) / np.linalg.norm(f_o.grad.numpy()) <= 1e-2
    for a, b in zip(gW, W_o):
        assert np.linalg.norm(a - b.grad.numpy()) / np.linalg.norm(b.grad.numpy()) <= 1e-2, a.shape
    # fp16 feature stream in / out: identical rounding points
    g_feat_h = np.zeros((M, 3 * C), np.float16)
    gW_h = [np.zeros(tuple(w.shape), np.float32) for w in W]
    kemu.call("tnl_mlp_backward", ctypes.byref(dims), packed, feat.numpy().astype(np.float16), 1, d.numpy(), M, None, gs.numpy(),
              grgb.numpy(), g_feat_h, *gW_h, None)
    assert np.array_equal(g_feat_h.astype(np.float32), g_feat)
    for a, b in zip(gW_h, gW):
        assert np.allclose(a, b, rtol=1e-5, atol=1e-5 * np.abs(b).max())
    # the 128-wide heads have no fused backward in this round: the ABI says so instead of launching
    from trinerflet_b200._lib import MlpDims
    big = MlpDims(3 * C, 128, 128)
    rc = kemu.lib().tnl_mlp_backward(ctypes.byref(big), kemu.p(packed), kemu.p(feat.numpy()), 0, kemu.p(d.numpy()), M, None,
                                     kemu.p(gs.numpy()), kemu.p(grgb.numpy()), kemu.p(g_feat), *[kemu.p(x) for x in gW], None)
    assert rc == -2


# ------------------------------------------------------------------------------------------------ marcher: other geometries
@pytest.mark.parametrize("bound,Hg,dt_gamma,max_steps", [(1.0, 128, 0.0, 256), (4.0, 128, 1.0 / 256, 512), (3.0, 32, 1.0 / 128, 128)])
def test_march_train_other_bounds_cascades_and_grid_sizes(bound, Hg, dt_gamma, max_steps):
    """the reference commands all use bound 1.5 (cascade 2, 128^3); the marcher's level selection / cone stepping for one, three
    cascades and a coarser grid stay bit-exact with the C oracle as well"""
    import math
    from trinerflet_b200 import scene as sc
    cas = 1 + math.ceil(math.log2(bound)) if bound > 1 else 1
    o, d = synthetic_rays(400, seed=int(bound * 10))
    o = (o * bound / 1.5).astype(np.float32)
    grid = sc.ball_density_grid(bound, 0.6 * bound, 1.0, Hg).numpy()
    rng = np.random.default_rng(1)
    grid = np.where(rng.random(grid.shape) < 0.02, 1 - grid, grid).astype(np.float32)
    bits = np.packbits(grid.reshape(-1) > 0.5, bitorder='little')
    aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    N = len(o)
    nears, fars = np.empty(N, np.float32), np.empty(N, np.float32)
    kemu.call("tnl_near_far_from_aabb", o, d, aabb, N, 0.2, nears, fars, None)
    n_o, f_o = orc.near_far_from_aabb(o, d, aabb, 0.2)
    assert _bits_equal(nears, n_o) and _bits_equal(fars, f_o)
    noises = rng.random(N).astype(np.float32)
    M = N * max_steps
    xyzs, dirs, deltas = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32)
    rays, counter = np.empty((N, 3), np.int32), np.zeros(2, np.int32)
    wsz = kemu.lib().tnl_march_rays_train_workspace(N)
    ws = np.zeros(wsz, np.uint8)
    kemu.call("tnl_march_rays_train", o, d, bits, bound, dt_gamma, max_steps, N, cas, Hg, M, nears, fars, xyzs, dirs, deltas, rays, counter,
              noises, ws, wsz, None)
    x_o, d_o, l_o, r_o, c_o = orc.march_rays_train(o, d, bound, bits, cas, Hg, nears, fars, noises, M, dt_gamma, max_steps)
    assert counter[0] > N and np.array_equal(counter, c_o) and np.array_equal(rays, r_o)
    assert _bits_equal(xyzs, x_o) and _bits_equal(deltas, l_o)




# ------------------------------------------------------------------------------------------------ peer-memory gradient exchange
def test_peer_memory_allreduce_kernels_p2p_variant():
    """csrc/tiles.cu tnl_tiles_allreduce / tnl_flat_allreduce, peer (non-multicast) variant, on the host build: the `peers` are three
    ordinary buffers standing in for three ranks' mappings of the symmetric plane-gradient buffer.  After every rank has run its
    share (tile k is owned by rank k % world) all buffers hold the average on the listed tiles and are untouched elsewhere."""
    rng = np.random.default_rng(3)
    world, R, C, T = 3, 128, 16, 32
    nt = R // T
    bufs = [rng.standard_normal((3, R, R, C)).astype(np.float32) for _ in range(world)]
    orig = [b.copy() for b in bufs]
    ids = np.sort(rng.permutation(3 * nt * nt)[:17]).astype(np.int32)
    cap = 32
    lst = np.zeros(cap, np.int32); lst[:17] = ids
    cnt = np.array([17], np.int32)
    peers = (ctypes.c_void_p * world)(*[b.ctypes.data for b in bufs])
    for r in range(world):
        rc = kemu.lib().tnl_tiles_allreduce(None, peers, kemu.p(lst), kemu.p(cnt), cap, R, C, T, r, world, ctypes.c_float(1.0 / world), None)
        assert rc == 0
    mean = sum(orig) / np.float32(world)
    mask = np.zeros((3, R, R), bool)
    for i in ids:
        p, ty, tx = i // (nt * nt), (i // nt) % nt, i % nt
        mask[p, ty * T:(ty + 1) * T, tx * T:(tx + 1) * T] = True
    for b, o in zip(bufs, orig):
        assert np.allclose(b[mask], mean[mask], rtol=1e-6, atol=1e-7)
        assert np.array_equal(b[~mask], o[~mask])
    assert np.array_equal(bufs[0][mask], bufs[1][mask]) and np.array_equal(bufs[0][mask], bufs[2][mask])    # bit-identical replicas
    # flat buffer (the MLP weight gradients), length a multiple of 4 * world
    flats = [rng.standard_normal(4 * world * 37).astype(np.float32) for _ in range(world)]
    want = sum(flats) / np.float32(world)
    fp = (ctypes.c_void_p * world)(*[f.ctypes.data for f in flats])
    for r in range(world):
        assert kemu.lib().tnl_flat_allreduce(None, fp, flats[0].size, r, world, ctypes.c_float(1.0 / world), None) == 0
    for f in flats:
        assert np.allclose(f, want, rtol=1e-6, atol=1e-7)
    assert kemu.lib().tnl_flat_allreduce(None, fp, 6, 0, world, ctypes.c_float(1.0), None) == -1        # length % 4
    assert kemu.lib().tnl_tiles_allreduce(None, None, kemu.p(lst), kemu.p(cnt), cap, R, C, T, 0, world, ctypes.c_float(1.0), None) == -1
