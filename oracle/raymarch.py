"""TEST INFRASTRUCTURE ONLY (oracle) -- ctypes/numpy front end of oracle/raymarch.c with the argument
order of the reference's Python wrappers (aux_libs/raymarching/raymarching.py:19-372)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_raymarch.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "all"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "raymarch.c")):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


u32, f32 = ctypes.c_uint32, ctypes.c_float


def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    rays_o, rays_d, aabb = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3), _f32(aabb)
    N = rays_o.shape[0]
    nears, fars = np.empty(N, np.float32), np.empty(N, np.float32)
    lib().orc_near_far_from_aabb(_p(rays_o), _p(rays_d), _p(aabb), u32(N), f32(min_near), _p(nears), _p(fars))
    return nears, fars


def sph_from_ray(rays_o, rays_d, radius):
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    N = rays_o.shape[0]
    coords = np.empty((N, 2), np.float32)
    lib().orc_sph_from_ray(_p(rays_o), _p(rays_d), f32(radius), u32(N), _p(coords))
    return coords


def morton3D(coords):
    coords = np.ascontiguousarray(coords, dtype=np.int32)
    N = coords.shape[0]
    out = np.empty(N, np.int32)
    lib().orc_morton3d(_p(coords), u32(N), _p(out))
    return out


def morton3D_invert(indices):
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    N = indices.shape[0]
    out = np.empty((N, 3), np.int32)
    lib().orc_morton3d_invert(_p(indices), u32(N), _p(out))
    return out


def packbits(grid, thresh):
    grid = _f32(grid)
    N = grid.size // 8
    out = np.empty(N, np.uint8)
    lib().orc_packbits(_p(grid), u32(N), f32(thresh), _p(out))
    return out


def march_rays_train(rays_o, rays_d, bound, bitfield, C, H, nears, fars, noises, M, dt_gamma=0.0, max_steps=1024):
    """Returns xyzs [M,3], dirs [M,3], deltas [M,2] (zero-initialised), rays [N,3] (ray order), counter [2]."""
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    nears, fars, noises = _f32(nears), _f32(fars), _f32(noises)
    bitfield = np.ascontiguousarray(bitfield, dtype=np.uint8)
    N = rays_o.shape[0]
    xyzs, dirs, deltas = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32)
    rays, counter = np.empty((N, 3), np.int32), np.zeros(2, np.int32)
    lib().orc_march_rays_train(_p(rays_o), _p(rays_d), _p(bitfield), f32(bound), f32(dt_gamma), u32(max_steps),
                               u32(N), u32(C), u32(H), u32(M), _p(nears), _p(fars), _p(xyzs), _p(dirs), _p(deltas),
                               _p(rays), _p(counter), _p(noises))
    return xyzs, dirs, deltas, rays, counter


def composite_rays_train_forward(sigmas, rgbs, deltas, rays, T_thresh=1e-4):
    sigmas, rgbs, deltas = _f32(sigmas), _f32(rgbs), _f32(deltas)
    rays = np.ascontiguousarray(rays, dtype=np.int32)
    M, N = sigmas.shape[0], rays.shape[0]
    ws, depth, image = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, 3), np.float32)
    lib().orc_composite_train_fwd(_p(sigmas), _p(rgbs), _p(deltas), _p(rays), u32(M), u32(N), f32(T_thresh),
                                  _p(ws), _p(depth), _p(image))
    return ws, depth, image


def composite_rays_train_backward(grad_ws, grad_image, sigmas, rgbs, deltas, rays, ws, image, T_thresh=1e-4):
    grad_ws, grad_image, sigmas, rgbs, deltas = map(_f32, (grad_ws, grad_image, sigmas, rgbs, deltas))
    ws, image = _f32(ws), _f32(image)
    rays = np.ascontiguousarray(rays, dtype=np.int32)
    M, N = sigmas.shape[0], rays.shape[0]
    gs, gc = np.zeros(M, np.float32), np.zeros((M, 3), np.float32)
    lib().orc_composite_train_bwd(_p(grad_ws), _p(grad_image), _p(sigmas), _p(rgbs), _p(deltas), _p(rays), _p(ws),
                                  _p(image), u32(M), u32(N), f32(T_thresh), _p(gs), _p(gc))
    return gs, gc


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, C, H, nears, fars, noises,
               align=-1, dt_gamma=0.0, max_steps=1024):
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    rays_alive = np.ascontiguousarray(rays_alive, dtype=np.int32)
    rays_t, nears, fars, noises = _f32(rays_t), _f32(nears), _f32(fars), _f32(noises)
    bitfield = np.ascontiguousarray(bitfield, dtype=np.uint8)
    M = n_alive * n_step
    if align > 0:
        M += align - (M % align)
    xyzs, dirs, deltas = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32)
    lib().orc_march_rays(u32(n_alive), u32(n_step), _p(rays_alive), _p(rays_t), _p(rays_o), _p(rays_d), f32(bound),
                         f32(dt_gamma), u32(max_steps), u32(C), u32(H), _p(bitfield), _p(nears), _p(fars),
                         _p(xyzs), _p(dirs), _p(deltas), _p(noises))
    return xyzs, dirs, deltas


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
    """In place on rays_alive, rays_t, weights_sum, depth, image (all must be contiguous numpy arrays)."""
    sigmas, rgbs, deltas = _f32(sigmas), _f32(rgbs), _f32(deltas)
    for a, dt in ((rays_alive, np.int32), (rays_t, np.float32), (weights_sum, np.float32), (depth, np.float32), (image, np.float32)):
        assert a.dtype == dt and a.flags["C_CONTIGUOUS"]
    lib().orc_composite_rays(u32(n_alive), u32(n_step), f32(T_thresh), _p(rays_alive), _p(rays_t), _p(sigmas), _p(rgbs),
                             _p(deltas), _p(weights_sum), _p(depth), _p(image))
