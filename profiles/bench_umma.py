"""Tensor-pipe cost per tcgen05.mma for the shapes / operand layouts the fused MLP uses (un-swizzled tiles)."""
import sys
import torch
sys.path.insert(0, ".")
from trinerflet_b200._lib import call, ptr, stream

out = torch.zeros(2, dtype=torch.int64, device="cuda")
print("shape                              layout     cycles/MMA (complete)  cycles/MMA (issue)")
for name, a_rows, b_rows, a_mn, b_mn, M, N, K in [
    ("fwd  M128 N64  (act x W)", 128, 64, 0, 0, 128, 64, 64),
    ("fwd  M128 N16  (act x W2)", 128, 16, 0, 0, 128, 16, 64),
    ("dX   M128 N64  (dOut x W mn)", 128, 64, 0, 1, 128, 64, 64),
    ("dX   M128 N96  (dh1 x W1 mn)", 128, 64, 0, 1, 128, 96, 64),
    ("dW   M64  N96  (mn x mn)", 128, 128, 1, 1, 64, 96, 128),
    ("dW   M64  N64  (mn x mn)", 128, 128, 1, 1, 64, 64, 128),
    ("dW   M64  N16  (mn x mn)", 128, 128, 1, 1, 64, 16, 128),
    ("dW   M128 N96  (mn x mn)", 128, 128, 1, 1, 128, 96, 128),
    ("dW   M128 N64  (mn x mn)", 128, 128, 1, 1, 128, 64, 128),
    ("M128 N128 k x k", 128, 128, 0, 0, 128, 128, 64),
    ("M128 N256 k x k", 128, 256, 0, 0, 128, 256, 64),
]:
    for reps in (1, 16):
        call("tnl_umma_bench", a_rows, b_rows, a_mn, b_mn, M, N, K, reps, ptr(out), stream())
        torch.cuda.synchronize()
        n = reps * K // 16
        c = out.cpu().tolist()
        print(f"{name:34s} reps={reps:3d} n={n:4d}   {c[0] / n:8.1f}   {c[1] / n:8.1f}   total {c[0]}")

print()
print("converged-warp issue, descriptor variants (cycles per MMA: complete / issue)")
for name, a, b, a_mn, b_mn, M, N, ks in [
    # (lbo, sbo, step, type)
    ("k-major none  M128 N64", (128, 8, 256, 0), (64, 8, 128, 0), 0, 0, 128, 64, 4),
    ("k-major none  M128 N16", (128, 8, 256, 0), (16, 8, 32, 0), 0, 0, 128, 16, 4),
    ("k-major none  M128 N128", (128, 8, 256, 0), (128, 8, 256, 0), 0, 0, 128, 128, 4),
    ("k-major sw128 M128 N64", (1, 64, 2, 2), (1, 64, 2, 2), 0, 0, 128, 64, 4),
    ("k-major sw128 M128 N16", (1, 64, 2, 2), (1, 64, 2, 2), 0, 0, 128, 16, 4),
    ("k-major sw128 M128 N128", (1, 64, 2, 2), (1, 64, 2, 2), 0, 0, 128, 128, 4),
    ("k-major sw128 M128 N256", (1, 64, 2, 2), (1, 64, 2, 2), 0, 0, 128, 256, 4),
    ("mn-major none  M64 N64", (8, 128, 16, 0), (8, 128, 16, 0), 1, 1, 64, 64, 8),
    ("mn-major none  M64 N96", (8, 128, 16, 0), (8, 128, 16, 0), 1, 1, 64, 96, 8),
    ("mn-major none  M128 N64", (8, 128, 16, 0), (8, 128, 16, 0), 1, 1, 128, 64, 8),
    ("mn-major sw128 M64 N64", (1, 64, 128, 2), (1, 64, 128, 2), 1, 1, 64, 64, 8),
    ("mn-major sw128 M128 N64", (512, 64, 128, 2), (1, 64, 128, 2), 1, 1, 128, 64, 8),
    ("mn-major sw128 M128 N128", (512, 64, 128, 2), (512, 64, 128, 2), 1, 1, 128, 128, 8),
]:
    for reps in (1, 32):
        call("tnl_umma_bench2", *a, *b, a_mn, b_mn, M, N, ks, reps, ptr(out), stream())
        torch.cuda.synchronize()
        n = reps * ks
        c = out.cpu().tolist()
        print(f"{name:28s} reps={reps:3d} n={n:4d}   {c[0] / n:8.1f}   {c[1] / n:8.1f}")
