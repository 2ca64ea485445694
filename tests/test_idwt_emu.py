"""CPU: the IDWT kernels' per-thread code (trinerflet_b200/csrc/idwt_core.cuh) executed by the host lock-step
emulator (tests/emu/idwt_emu.cpp) against the oracle -- checks tiling, halo, ring-buffer and zero-padding index
arithmetic for every channel-chunk variant without a GPU.  (The GPU tests check the real kernels.)"""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import wavelet as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tests", "emu", "_build", "libidwt_emu.so")


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(ROOT, "tests", "emu", "idwt_emu.cpp")
    core = os.path.join(ROOT, "trinerflet_b200", "csrc", "idwt_core.cuh")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(core)):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SO, src])
    return ctypes.CDLL(SO)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("C,n", [(8, 8), (16, 32), (32, 64), (24, 40), (16, 104), (48, 16)])
def test_emulated_kernels_match_oracle(emu, C, n):
    g = torch.Generator().manual_seed(C * 1000 + n)
    x = torch.randn(3, C, n, n, generator=g, requires_grad=True)
    yh = torch.randn(3, C, 3, n, n, generator=g, requires_grad=True)
    ref = W.build_planes(x, [yh])
    gout = torch.randn(ref.shape, generator=g)
    ref.backward(gout)
    xc = np.ascontiguousarray(x.detach().permute(0, 2, 3, 1).numpy())
    yc = np.ascontiguousarray(yh.detach().permute(0, 2, 3, 4, 1).numpy())
    out = np.full((3, 2 * n, 2 * n, C), np.nan, np.float32)
    asum = np.zeros(1, np.float32)
    emu.emu_idwt_level_forward(_p(xc), _p(yc), _p(out), ctypes.c_uint32(n), ctypes.c_uint32(C), _p(asum))
    assert abs(asum[0] - yh.detach().abs().sum().item()) <= 1e-4 * asum[0]     # every coefficient counted exactly once
    scale = ref.abs().max().item()
    assert np.isfinite(out).all()
    assert (torch.from_numpy(out).permute(0, 3, 1, 2) - ref.detach()).abs().max().item() <= 1e-5 * scale
    gc = np.ascontiguousarray(gout.permute(0, 2, 3, 1).numpy())
    gx = np.full((3, n, n, C), np.nan, np.float32)
    gyh = np.full((3, 3, n, n, C), np.nan, np.float32)
    emu.emu_idwt_level_backward(_p(gc), _p(gx), _p(gyh), ctypes.c_uint32(n), ctypes.c_uint32(C), None, ctypes.c_float(0))
    assert (torch.from_numpy(gx).permute(0, 3, 1, 2) - x.grad).abs().max().item() <= 1e-5 * x.grad.abs().max().item()
    assert (torch.from_numpy(gyh).permute(0, 4, 1, 2, 3) - yh.grad).abs().max().item() <= 1e-5 * yh.grad.abs().max().item()
    gyh2 = np.full((3, 3, n, n, C), np.nan, np.float32)      # fused L1-regulariser gradient: + reg * sign(yh)
    emu.emu_idwt_level_backward(_p(gc), _p(gx), _p(gyh2), ctypes.c_uint32(n), ctypes.c_uint32(C), _p(yc), ctypes.c_float(0.25))
    want = yh.grad + 0.25 * torch.sign(yh.detach())
    assert (torch.from_numpy(gyh2).permute(0, 4, 1, 2, 3) - want).abs().max().item() <= 1e-5 * want.abs().max().item()


@pytest.mark.parametrize("C,n", [(16, 32), (24, 48), (8, 64)])
def test_emulated_worklist_kernels_match_oracle_on_their_blocks(emu, C, n):
    """Work-list mode of the level kernels (idwt_geom_item + the run lists of trinerflet_b200/idwt_plan.py) on the host
    emulator: active blocks reproduce the oracle on exactly their outputs, nothing else is written; the adjoint, fed a
    gradient confined to the marked tiles, matches the oracle on the blocks whose halo can see that gradient."""
    from trinerflet_b200.idwt_plan import _items, level_maps
    g = torch.Generator().manual_seed(C * 77 + n)
    nb = n // 16
    flags = torch.rand(3, nb, nb, generator=g) < 0.4
    flags[:, 0, 0] = True
    flags[:, nb - 1, nb - 1] = False
    fwd_maps, bwd_maps = level_maps(flags, 1)
    mf, mb = fwd_maps[0].numpy(), bwd_maps[0].numpy()
    x = torch.randn(3, C, n, n, generator=g, requires_grad=True)
    yh = torch.randn(3, C, 3, n, n, generator=g, requires_grad=True)
    ref = W.build_planes(x, [yh])
    tile_mask = flags.repeat_interleave(32, 1).repeat_interleave(32, 2)[:, None]           # [3,1,2n,2n] output pixels
    gout = torch.randn(ref.shape, generator=g) * tile_mask
    ref.backward(gout)
    xc = np.ascontiguousarray(x.detach().permute(0, 2, 3, 1).numpy())
    yc = np.ascontiguousarray(yh.detach().permute(0, 2, 3, 4, 1).numpy())
    # ---- forward: active items, runs of at most 2 blocks
    items = np.ascontiguousarray(_items(mf, True, 2))
    out = np.full((3, 2 * n, 2 * n, C), np.nan, np.float32)
    asum = np.zeros(1, np.float32)
    emu.emu_idwt_level_forward_items(_p(xc), _p(yc), _p(out), ctypes.c_uint32(n), ctypes.c_uint32(C), _p(asum), _p(items),
                                     ctypes.c_int32(items.shape[0]))
    out_t = torch.from_numpy(out).permute(0, 3, 1, 2)
    m = tile_mask.expand_as(out_t)
    assert torch.isnan(out_t[~m]).all()                                    # clean blocks: untouched
    assert (out_t[m] - ref.detach()[m]).abs().max().item() <= 1e-5 * ref.abs().max().item()
    coef_mask = torch.from_numpy(mf).repeat_interleave(16, 1).repeat_interleave(16, 2)[:, None, None]     # [3,1,1,n,n]
    want_abs = (yh.detach().abs() * coef_mask).sum().item()
    assert abs(asum[0] - want_abs) <= 1e-4 * max(want_abs, 1.0)           # |yh| of exactly the active blocks
    # ---- backward: blocks that can see the gradient (bwd map); regulariser gradient included
    items_b = np.ascontiguousarray(_items(mb, True, 3))
    gc = np.ascontiguousarray(gout.permute(0, 2, 3, 1).numpy())
    gx = np.full((3, n, n, C), np.nan, np.float32)
    gyh = np.full((3, 3, n, n, C), np.nan, np.float32)
    emu.emu_idwt_level_backward_items(_p(gc), _p(gx), _p(gyh), ctypes.c_uint32(n), ctypes.c_uint32(C), _p(yc), ctypes.c_float(0.25),
                                      _p(items_b), ctypes.c_int32(items_b.shape[0]))
    act = torch.from_numpy(mb).repeat_interleave(16, 1).repeat_interleave(16, 2)           # [3,n,n] coefficients
    gx_t = torch.from_numpy(gx).permute(0, 3, 1, 2)
    gyh_t = torch.from_numpy(gyh).permute(0, 4, 1, 2, 3)
    ax = act[:, None].expand_as(gx_t)
    ay = act[:, None, None].expand_as(gyh_t)
    assert torch.isnan(gx_t[~ax]).all() and torch.isnan(gyh_t[~ay]).all()
    assert (gx_t[ax] - x.grad[ax]).abs().max().item() <= 1e-5 * x.grad.abs().max().item()
    want = yh.grad + 0.25 * torch.sign(yh.detach())
    assert (gyh_t[ay] - want[ay]).abs().max().item() <= 1e-5 * want.abs().max().item()
    # outside the active blocks the true gradients are exactly zero (that is what the clean pass writes)
    assert float(x.grad[~ax].abs().sum()) == 0.0 and float(yh.grad[~ay].abs().sum()) == 0.0
