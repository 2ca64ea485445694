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
