"""Generates tests/golden/raymarch_ref.npz from the UNMODIFIED reference CUDA kernels (oracle/_ref/_raymarching_ref.so,
compiled by oracle/build_ref.py from /root/reference/aux_libs/raymarching/src).  Needs a GPU:

    gpurun -- 'python tests/golden/make_raymarch_golden.py gpurun_out/raymarch_ref.npz'   # then copy into tests/golden/

The fixture pins the CPU oracle (oracle/raymarch.c) in the `-m "not gpu"` suite: tests/test_oracle_raymarch.py."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402
from tests.util import random_bitfield, synthetic_rays  # noqa: E402


def main(out):
    ref = build_ref.load_ref("raymarching")
    assert ref is not None, "oracle/_ref/_raymarching_ref.so missing"
    N, max_steps, dt_gamma, bound, C, H = 192, 256, 0.0, 1.5, 2, 128
    o, d = synthetic_rays(N, seed=7)
    grid, bits = random_bitfield(seed=7)
    ro, rd, bf = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), bits.cuda()
    aabb = torch.tensor([-bound] * 3 + [bound] * 3, device="cuda")
    nears, fars = torch.empty(N, device="cuda"), torch.empty(N, device="cuda")
    ref.near_far_from_aabb(ro, rd, aabb, N, 0.2, nears, fars)
    torch.manual_seed(0)
    noises = torch.rand(N, device="cuda")
    M = N * 96
    xyzs, dirs, deltas = torch.zeros(M, 3, device="cuda"), torch.zeros(M, 3, device="cuda"), torch.zeros(M, 2, device="cuda")
    rays = torch.empty(N, 3, dtype=torch.int32, device="cuda")
    counter = torch.zeros(2, dtype=torch.int32, device="cuda")
    ref.march_rays_train(ro, rd, bf, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs, deltas, rays, counter, noises)
    total = int(counter[0])
    assert 0 < total <= M, total
    sig = torch.rand(M, device="cuda") * 40
    rgb = torch.rand(M, 3, device="cuda")
    ws, dp, im = torch.empty(N, device="cuda"), torch.empty(N, device="cuda"), torch.empty(N, 3, device="cuda")
    ref.composite_rays_train_forward(sig, rgb, deltas, rays, M, N, 1e-4, ws, dp, im)
    g = torch.rand(2, 4096, device="cuda")
    packed = torch.empty(2 * 4096 // 8, dtype=torch.uint8, device="cuda")
    ref.packbits(g, packed.numel(), 0.37, packed)
    coords = torch.randint(0, 128, (512, 3), dtype=torch.int32, device="cuda")
    mort = torch.empty(512, dtype=torch.int32, device="cuda")
    ref.morton3D(coords, 512, mort)
    keep = total + 8
    np.savez_compressed(out, rays_o=o, rays_d=d, bitfield=bits.numpy(), noises=noises.cpu().numpy(), nears=nears.cpu().numpy(),
                        fars=fars.cpu().numpy(), M=M, dt_gamma=dt_gamma, max_steps=max_steps, counter=counter.cpu().numpy(),
                        rays=rays.cpu().numpy(), xyzs=xyzs.cpu().numpy()[:keep], deltas=deltas.cpu().numpy()[:keep],
                        sigmas=sig.cpu().numpy()[:keep], rgbs=rgb.cpu().numpy()[:keep], weights_sum=ws.cpu().numpy(),
                        image=im.cpu().numpy(), grid=g.cpu().numpy(), thresh=0.37, packed=packed.cpu().numpy(),
                        coords=coords.cpu().numpy(), morton=mort.cpu().numpy())
    print("wrote", out, "samples", total)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "raymarch_ref.npz"))
