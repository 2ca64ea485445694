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
    t = timeit(lambda: call("tnl_idwt_level_backward", ptr(out), ptr(gx), ptr(gyh), n, C, None, None, 0.0, stream()))
    print(f"idwt_bwd  n={n} C={C}: {t:.3f} ms  {by / t / 1e6:.0f} GB/s algorithmic")
    t = timeit(lambda: call("tnl_idwt_level_backward", ptr(out), ptr(gx), ptr(gyh), n, C, ptr(yh), ptr(g1), 1.0, stream()))
    print(f"idwt_bwd+reg n={n} C={C}: {t:.3f} ms  {by / t / 1e6:.0f} GB/s algorithmic")

if __name__ == "__main__":
    which = sys.argv[1:] or ["idwt"]
    if "idwt" in which:
        bench_idwt(32, 1024); bench_idwt(32, 512); bench_idwt(16, 512)
