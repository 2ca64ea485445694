"""Shared helpers for the GPU parity tests (channels-last conversion, tolerances)."""
import numpy as np
import torch


def cl_planes(t):
    """logical [3,C,n,n] (any strides) -> channels-last strided tensor with the same logical shape"""
    return t.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)


def cl_coefs(t):
    """logical [3,C,3,n,n] -> stored [3][3][n][n][C]"""
    return t.permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3)


def rel_linf(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def synthetic_rays(n, seed=0, bound=1.5):
    """rays from a ring of cameras at radius ~4 looking roughly at the origin (CPU fp32 numpy)."""
    rng = np.random.default_rng(seed)
    o = rng.normal(size=(n, 3))
    o = 4.03 * o / np.linalg.norm(o, axis=1, keepdims=True)
    tgt = rng.uniform(-0.9, 0.9, size=(n, 3))
    d = tgt - o
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    return o.astype(np.float32), d.astype(np.float32)


def random_bitfield(seed=0, cascade=2, H=128, radius=0.8, bound=1.5):
    """occupancy: noisy ball, different per cascade (uint8 [cascade*H^3/8], Morton order)"""
    from trinerflet_b200 import scene
    grid = scene.ball_density_grid(bound, radius, 1.0, H).numpy()
    rng = np.random.default_rng(seed)
    flip = rng.random(grid.shape) < 0.02
    grid = np.where(flip, 1.0 - grid, grid).astype(np.float32)
    bits = np.packbits(grid.reshape(-1) > 0.5, bitorder='little')
    return torch.from_numpy(grid), torch.from_numpy(bits)


def build_emu(name, deps=(), flags=()):
    """Compile tests/emu/<name>.cpp (the host emulator of a kernel's per-thread code) into tests/emu/_build/lib<name>.so
    if it is older than its sources, and load it with ctypes.  TEST-ONLY code, never part of the product library."""
    import ctypes
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "tests", "emu", name + ".cpp")
    so = os.path.join(root, "tests", "emu", "_build", "lib" + name + ".so")
    newest = max(os.path.getmtime(p) for p in [src] + [os.path.join(root, d) for d in deps])
    if not os.path.exists(so) or os.path.getmtime(so) < newest:
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", *flags, "-o", so, src])
    return ctypes.CDLL(so)
