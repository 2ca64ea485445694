"""Pins the tcgen05 descriptor conventions of csrc/umma.cuh on real hardware: runs single products through
tnl_umma_probe and compares the raw TMEM dump with numpy.  For M = 64 it also searches where each row landed."""
import sys
import numpy as np
import torch

sys.path.insert(0, ".")
from trinerflet_b200._lib import call, ptr, stream  # noqa: E402


def run(M, N, K, a_mn, b_mn, ncols=None, seed=0):
    g = torch.Generator().manual_seed(seed)
    A = (torch.randint(-4, 5, (M, K), generator=g).float() / 4).half()      # exactly representable, exact products
    B = (torch.randint(-4, 5, (N, K), generator=g).float() / 4).half()
    D = A.float() @ B.float().t()
    As = A.t().contiguous() if a_mn else A.contiguous()
    Bs = B.t().contiguous() if b_mn else B.contiguous()
    ncols = ncols or max(8, N)
    out = torch.full((128, ncols), float("nan"), device="cuda")
    Ad, Bd = As.cuda(), Bs.cuda()
    call("tnl_umma_probe", ptr(Ad), As.shape[0], As.shape[1], ptr(Bd), Bs.shape[0], Bs.shape[1], int(a_mn), int(b_mn), M, N, K,
         ptr(out), ncols, stream())
    torch.cuda.synchronize()
    return D.numpy(), out.cpu().numpy()


def report(name, M, N, K, a_mn, b_mn, ncols=None):
    D, out = run(M, N, K, a_mn, b_mn, ncols)
    if M == 128:
        err = np.abs(out[:, :N] - D).max()
        print(f"{name}: M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn}  max|err| vs lane=row,col=col: {err}")
        return err == 0
    # M = 64: hypothesis row i -> lane (i // 16) * 32 + i % 16
    lanes = np.array([(i // 16) * 32 + i % 16 for i in range(64)])
    err = np.abs(out[lanes, :N] - D).max()
    print(f"{name}: M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn}  max|err| vs lane=(i//16)*32+i%16: {err}")
    if err != 0:
        for i in range(0, 64, 5):
            hits = [(l, c) for l in range(128) for c in range(out.shape[1] - N + 1) if np.array_equal(out[l, c:c + N], D[i])]
            print("   row", i, "found at (lane, col0):", hits[:4])
    return err == 0


if __name__ == "__main__":
    ok = True
    ok &= report("fwd L1 ", 128, 64, 96, False, False)
    ok &= report("fwd L2 ", 128, 16, 64, False, False)
    ok &= report("fwd L5n8", 128, 16, 64, False, False)
    ok &= report("dX W5  ", 128, 64, 16, False, True)
    ok &= report("dX W1  ", 128, 96, 64, False, True)
    ok &= report("dW1    ", 64, 96, 128, True, True, ncols=128)
    ok &= report("dW4    ", 64, 64, 128, True, True, ncols=128)
    ok &= report("dW5t   ", 64, 16, 128, True, True, ncols=64)
    print("ALL OK" if ok else "MISMATCH")
