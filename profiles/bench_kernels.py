"""Micro-benchmarks of single ABI calls at base-light sizes (CUDA events, L2-exceeding inputs). Usage:
   python profiles/bench_kernels.py idwt|sample|mlp|march|composite ..."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from trinerflet_b200 import _lib, scene
from trinerflet_b200._lib import call, ptr, stream
from trinerflet_b200.triplane_encoder import cl_empty_planes, cl_empty_coefs

def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def bench_idwt(C=32, n=1024):
    x = cl_empty_planes(C, n, device="cuda").normal_(); yh = cl_empty_coefs(C, n, device="cuda").normal_()
    out = cl_empty_planes(C, 2 * n, device="cuda"); asum = torch.zeros(1, device="cuda")
    gx = cl_empty_planes(C, n, device="cuda"); gyh = cl_empty_coefs(C, n, device="cuda"); g1 = torch.ones(1, device="cuda")
    by = 2 * 3 * C * (2 * n) ** 2 * 4
    t = timeit(lambda: call("tnl_idwt_level_forward", ptr(x), ptr(yh), ptr(out), n, C, ptr(asum), stream()))
    print(f"idwt_fwd  n={n} C={C}: {t:.3f} ms  {by / t / 1e6:.0f} GB/s algorithmic")
    t = timeit(lambda: call("tnl_idwt_level_backward", ptr(out), ptr(gx), ptr(gyh), n, C, None, None, 0.0, 0, 3, stream()))
    print(f"idwt_bwd  n={n} C={C}: {t:.3f} ms  {by / t / 1e6:.0f} GB/s algorithmic")
    t = timeit(lambda: call("tnl_idwt_level_backward", ptr(out), ptr(gx), ptr(gyh), n, C, ptr(yh), ptr(g1), 1.0, 0, 3, stream()))
    print(f"idwt_bwd+reg n={n} C={C}: {t:.3f} ms  {by / t / 1e6:.0f} GB/s algorithmic")

if __name__ == "__main__":
    which = sys.argv[1:] or ["idwt"]
    if "idwt" in which:
        bench_idwt(32, 1024); bench_idwt(32, 512); bench_idwt(16, 512)

def bench_sample(C=32, R=2048, n_rays=60000):
    import numpy as np
    from trinerflet_b200 import raymarching as rm
    from trinerflet_b200.triplane_encoder import cell_sort
    sc = scene.make_scene()
    ro, rd, _ = scene.sample_batch(sc, n_rays, torch.Generator().manual_seed(0))
    ro, rd = ro.cuda(), rd.cuda()
    grid = scene.ball_density_grid(1.5, 0.75); bits = scene.packbits_cpu(grid, 0.5).cuda()
    aabb = torch.tensor([-1.5] * 3 + [1.5] * 3, device="cuda")
    nears, fars = rm.near_far_from_aabb(ro, rd, aabb, 0.2)
    cnt = torch.zeros(2, dtype=torch.int32, device="cuda")
    xyzs, dirs, deltas, rays = rm.march_rays_train(ro, rd, 1.5, bits, 2, 128, nears, fars, cnt, -1, True, 128, True, 0, 1024)
    M = xyzs.shape[0]; print("M", M)
    planes = cl_empty_planes(C, R, device="cuda").normal_()
    feat = torch.empty(M, 3 * C, device="cuda"); gfeat = torch.randn(M, 3 * C, device="cuda")
    gpl = cl_empty_planes(C, R, device="cuda", zero=True)
    inv = float(np.float32(1) / np.float32(1.5))
    for G in (0, 32, 64, 128):
        perm = None
        if G:
            t = timeit(lambda: cell_sort(xyzs, 1.5, None, G), iters=5)
            perm = cell_sort(xyzs, 1.5, None, G)
            print(f"cell_sort G={G}: {t:.3f} ms")
        tf = timeit(lambda: call("tnl_sample_planes_forward", ptr(planes), ptr(xyzs), M, R, C, inv, 1, None, ptr(perm), ptr(feat), 0, stream()))
        tb = timeit(lambda: call("tnl_sample_planes_backward", ptr(gfeat), 0, ptr(xyzs), M, R, C, inv, 1, None, ptr(perm), ptr(gpl), stream()))
        print(f"G={G}: sample_fwd {tf:.3f} ms  sample_bwd {tb:.3f} ms")

if __name__ == "__main__" and "sample" in sys.argv[1:]:
    bench_sample()
