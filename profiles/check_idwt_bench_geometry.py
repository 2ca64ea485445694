"""one-off: work-list IDWT == dense IDWT on the marked tiles at the bench geometry (n0=64, 5 levels, R=2048) with the
ball-occupancy tile flags, on the host build of the kernels"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import kemu
from tests.test_kernels_emu import _idwt_chain_emu, _idwt_chain_bwd_emu
from trinerflet_b200 import scene
from trinerflet_b200.idwt_plan import IdwtPlan
C, n0, L = 16, 64, 5
R = n0 * 2 ** L
T = R // 32
bits = scene.packbits_cpu(scene.ball_density_grid(1.5, 0.75, 1.0, 128), 0.5).numpy()
flags = np.zeros((3, T, T), np.uint8)
kemu.call("tnl_mark_dirty_tiles", bits, 2, 128, 1.5, R, 32, 2, flags, None)
print("tile fraction", flags.mean())
plan = IdwtPlan(R, n0, L, C, "cpu").update(torch.from_numpy(flags))
print(plan.stats)
g = torch.Generator().manual_seed(0)
pf = torch.randn(3, C, n0, n0, generator=g)
coefs = [0.1 * torch.randn(3, C, 3, n0 * 2 ** l, n0 * 2 ** l, generator=g) for l in range(L)]
t = time.time(); dense, abs_d = _idwt_chain_emu(pf, coefs); print("dense fwd", time.time() - t)
t = time.time(); sparse, abs_s = _idwt_chain_emu(pf, coefs, plan); print("sparse fwd", time.time() - t)
mask = np.repeat(np.repeat(flags.astype(bool), 32, axis=1), 32, axis=2)[..., None]
print("fwd equal on marked tiles:", np.array_equal(np.where(mask, sparse, 0).view(np.uint32), np.where(mask, dense, 0).view(np.uint32)))
print("abs sums rel err:", np.abs(abs_s - abs_d).max() / abs_d.max())
del sparse
gout = (np.random.default_rng(1).standard_normal(dense.shape).astype(np.float32) * mask).astype(np.float32)
del dense
t = time.time(); gx_d, gy_d = _idwt_chain_bwd_emu(gout, coefs, n0, 0.7); print("dense bwd", time.time() - t)
t = time.time(); gx_s, gy_s = _idwt_chain_bwd_emu(gout, coefs, n0, 0.7, plan); print("sparse bwd", time.time() - t)
print("bwd g_x equal:", np.array_equal(gx_s.view(np.uint32), gx_d.view(np.uint32)))
print("bwd g_yh equal:", [bool(np.array_equal(a.view(np.uint32), b.view(np.uint32))) for a, b in zip(gy_s, gy_d)])
