"""GPU parity: wavelet plane reconstruction (IDWT fwd / adjoint bwd) and tri-plane sampling against the oracle
(oracle/wavelet.py, oracle/field.py) and the reference-generated golden fixture tests/golden/encoder_fp32.npz.
Stated tolerances: planes and features rel-L_inf <= 1e-5 (fp32); gradients rel-L2 <= 1e-4 (fp32;
float atomics reorder the scatter sums)."""
import os

import numpy as np
import pytest
import torch

from tests.util import cl_coefs, cl_planes, rel_l2, rel_linf

pytestmark = pytest.mark.gpu
TOL_FWD = 1e-5
TOL_GRAD = 1e-4


def _rand_coefs(C, n0, levels, seed=0, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    pf = 0.1 * torch.randn(3, C, n0, n0, generator=g)
    coefs = [0.05 * torch.randn(3, C, 3, n0 * 2 ** l, n0 * 2 ** l, generator=g) for l in range(levels)]
    return pf.to(device), [c.to(device) for c in coefs]


@pytest.mark.parametrize("C,n0,levels", [(8, 8, 1), (16, 16, 2), (32, 8, 3), (48, 24, 1), (24, 40, 1)])
def test_build_planes_matches_oracle(C, n0, levels):
    from oracle import wavelet as ow
    from trinerflet_b200.triplane_encoder import build_planes
    pf, coefs = _rand_coefs(C, n0, levels)
    pf_o = pf.clone().requires_grad_(True)
    coefs_o = [c.clone().requires_grad_(True) for c in coefs]
    ref = ow.build_planes(pf_o, coefs_o)
    gout = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1))
    ref.backward(gout)
    pf_g = cl_planes(pf.cuda()).requires_grad_(True)
    coefs_g = [cl_coefs(c.cuda()).requires_grad_(True) for c in coefs]
    out = build_planes(pf_g, coefs_g)
    assert out.shape == ref.shape
    assert rel_linf(out, ref) <= TOL_FWD
    out.backward(gout.cuda())
    assert rel_l2(pf_g.grad, pf_o.grad) <= TOL_GRAD
    for a, b in zip(coefs_g, coefs_o):
        assert a.grad.shape == b.grad.shape and rel_l2(a.grad, b.grad) <= TOL_GRAD


def test_fused_wavelet_regulariser():
    """encoder.wavelet_l1 (value from the IDWT forward, gradient inside the IDWT backward) vs the reference's literal
    torch expression (nerf/utils.py:640-655)."""
    from trinerflet_b200 import trainer
    from trinerflet_b200.triplane_encoder import TriPlaneVolume
    enc = TriPlaneVolume(number_of_features=16, plane_resolution=256, inner_multi_res_scale=8).cuda()
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        for p in enc.planes_features_wavelet_coefs:
            p.copy_(0.05 * torch.randn(p.shape, generator=g))
            p[:, :, :, ::3, ::5] = 0.0     # exact zeros: sign(0) = 0
    w = torch.randn(3, 16, 256, 256, generator=g).cuda()
    grads = []
    for fused in (True, False):
        enc.zero_grad(set_to_none=True); enc.reset_cahce()
        planes = enc.get_planes()
        reg = trainer.wavelet_regulariser(enc, 0.2, fused=fused)
        ((planes * w).sum() * 1e-3 + 64.0 * reg).backward()
        grads.append((float(reg), [p.grad.clone() for p in enc.parameters()]))
    assert abs(grads[0][0] - grads[1][0]) <= 1e-5 * abs(grads[1][0])
    for a, b in zip(grads[0][1], grads[1][1]):
        assert rel_l2(a, b) <= 1e-5


def test_golden_encoder_fixture(golden_dir):
    """The reference's own TriPlaneVolume (run on CPU by tests/golden/make_golden.py) vs our module on the GPU.
    The fixture uses C = 4 < 8, so channels are zero-padded to 8 (independent channels: exact)."""
    from trinerflet_b200.triplane_encoder import TriPlaneVolume
    g = np.load(os.path.join(golden_dir, "encoder_fp32.npz"))
    C, R, S, bound = int(g["C"]), int(g["R"]), int(g["S"]), float(g["bound"])
    Cp = 8
    enc = TriPlaneVolume(number_of_features=Cp, plane_resolution=R, inner_multi_res_scale=S).cuda()
    with torch.no_grad():
        enc.planes_features.zero_()
        enc.planes_features[:, :C].copy_(torch.from_numpy(g["planes_features"]))
        for i, p in enumerate(enc.planes_features_wavelet_coefs):
            p.zero_()
            p[:, :C].copy_(torch.from_numpy(g[f"coef{i}"]))
    planes = enc.get_planes()
    assert rel_linf(planes[:, :C], torch.from_numpy(g["planes"])) <= TOL_FWD
    xyz = torch.from_numpy(g["xyz"]).cuda()
    feat = enc(xyz, bound).view(-1, 3, Cp)[:, :, :C].reshape(xyz.shape[0], -1)
    # the fixture was produced on CPU (true division by bound); CUDA uses x * (1/bound): <= 1 ulp of the coordinate
    assert rel_linf(feat, torch.from_numpy(g["feat"])) <= 5e-5
    (feat * torch.from_numpy(g["wfeat"]).cuda()).sum().backward()
    assert rel_l2(enc.planes_features.grad[:, :C], torch.from_numpy(g["g_planes_features"])) <= 5e-4
    for i, p in enumerate(enc.planes_features_wavelet_coefs):
        assert rel_l2(p.grad[:, :C], torch.from_numpy(g[f"g_coef{i}"])) <= 5e-4


@pytest.mark.parametrize("C,R,fp16", [(16, 64, False), (32, 256, False), (32, 256, True), (48, 128, True)])
def test_sampling_matches_oracle(C, R, fp16):
    from oracle import field as of
    from trinerflet_b200.triplane_encoder import sample_planes
    g = torch.Generator().manual_seed(5)
    planes = torch.randn(3, C, R, R, generator=g)
    M, bound = 20000, 1.5
    xyz = (torch.rand(M, 3, generator=g) * 2 - 1) * bound
    xyz[:6] = torch.tensor([[-1.5, -1.5, -1.5], [1.5, 1.5, 1.5], [0, 0, 0], [1.5, -1.5, 0.3], [1.7, -2.0, 0.1], [0.75, 0.75, 0.75]])
    w = torch.randn(M, 3 * C, generator=g)
    p_o = planes.clone().requires_grad_(True)
    f_o = of.sample_planes(p_o, xyz, bound, fp16=fp16, recip_mul=True)
    (f_o * w).sum().backward()
    p_g = cl_planes(planes.cuda()).requires_grad_(True)
    f_g = sample_planes(p_g, xyz.cuda(), bound, fp16_coords=fp16)
    assert rel_linf(f_g, f_o) <= TOL_FWD
    (f_g * w.cuda()).sum().backward()
    assert rel_l2(p_g.grad, p_o.grad) <= TOL_GRAD
    # spatially sorted visiting order: identical features, gradients equal up to atomic summation order
    from trinerflet_b200.triplane_encoder import cell_sort
    perm = cell_sort(xyz.cuda(), bound, None, 16)
    assert torch.equal(torch.sort(perm.long())[0], torch.arange(M, device="cuda"))
    p_s = cl_planes(planes.cuda()).requires_grad_(True)
    f_s = sample_planes(p_s, xyz.cuda(), bound, fp16_coords=fp16, perm=perm)
    assert torch.equal(f_s, f_g)
    (f_s * w.cuda()).sum().backward()
    assert rel_l2(p_s.grad, p_o.grad) <= TOL_GRAD
    # fp16 feature stream: the fp32 features rounded to nearest fp16, and the matching scatter of an fp16 gradient
    p_h = cl_planes(planes.cuda()).requires_grad_(True)
    f_h = sample_planes(p_h, xyz.cuda(), bound, fp16_coords=fp16, half_out=True)
    assert f_h.dtype == torch.float16 and torch.equal(f_h, f_g.half())
    (f_h.float() * w.cuda().half().float()).sum().backward()
    p_r = cl_planes(planes.cuda()).requires_grad_(True)
    (sample_planes(p_r, xyz.cuda(), bound, fp16_coords=fp16) * w.cuda().half().float()).sum().backward()
    assert rel_l2(p_h.grad, p_r.grad) <= TOL_GRAD
    # n_valid: rows past the counter are skipped (zeros out, no gradient)
    nv = torch.tensor([M // 2], dtype=torch.int32, device="cuda")
    p_g2 = cl_planes(planes.cuda()).requires_grad_(True)
    f_g2 = sample_planes(p_g2, xyz.cuda(), bound, fp16_coords=fp16, n_valid=nv)
    assert torch.equal(f_g2[:M // 2], f_g[:M // 2]) and float(f_g2[M // 2:].abs().sum()) == 0.0


def test_full_size_properties():
    """BASELINE sizes (C=32, 64 -> 2048, 5 levels), size-independent properties: DC gain, linearity, adjoint identity."""
    from trinerflet_b200.triplane_encoder import build_planes, cl_empty_coefs, cl_empty_planes
    C, n0, L = 32, 64, 5
    dev = "cuda"
    pf = cl_empty_planes(C, n0, device=dev); pf.fill_(0.37)
    coefs = [cl_empty_coefs(C, n0 * 2 ** l, device=dev, zero=True) for l in range(L)]
    planes = build_planes(pf, coefs)
    assert planes.shape == (3, C, 2048, 2048)
    inner = planes[:, :, 256:-256, 256:-256]
    assert (inner - 0.37).abs().max().item() <= 1e-5          # IDWT(2c, 0) = c away from the zero-padded border
    g = torch.Generator(device=dev).manual_seed(0)
    x1 = [cl_empty_planes(C, n0, device=dev).normal_(generator=g)] + [cl_empty_coefs(C, n0 * 2 ** l, device=dev).normal_(generator=g) for l in range(L)]
    x2 = [cl_empty_planes(C, n0, device=dev).normal_(generator=g)] + [cl_empty_coefs(C, n0 * 2 ** l, device=dev).normal_(generator=g) for l in range(L)]
    y1 = build_planes(x1[0], x1[1:]); y2 = build_planes(x2[0], x2[1:])
    y12 = build_planes(x1[0] + 2 * x2[0], [a + 2 * b for a, b in zip(x1[1:], x2[1:])])
    assert rel_linf(y12, y1 + 2 * y2) <= 1e-5                 # linearity
    del y12, y2
    xs = [t.clone().requires_grad_(True) for t in x1]
    y = build_planes(xs[0], xs[1:])
    gy = torch.empty_like(y).normal_(generator=g)
    y.backward(gy)
    lhs = (y.detach().double() * gy.double()).sum().item()     # <A x, g>
    rhs = sum((t.detach().double() * t.grad.double()).sum().item() for t in xs)   # <x, A^T g>
    assert abs(lhs - rhs) <= 1e-5 * abs(lhs)


def _plan_for(flags, R, n0, levels, C):
    from trinerflet_b200.idwt_plan import IdwtPlan
    return IdwtPlan(R, n0, levels, C, "cuda").update(flags.cuda())


@pytest.mark.parametrize("C,n0,levels,density", [(16, 64, 2, 0.05), (32, 32, 3, 0.02), (24, 16, 1, 0.3), (16, 16, 2, 0.0)])
def test_worklist_idwt_equals_dense_on_the_marked_tiles(C, n0, levels, density):
    """Work-list reconstruction (idwt_plan.py): bit-identical to the dense kernels inside the marked tiles, identical
    |yh| sums up to the order of the float atomics; the adjoint, fed a gradient that vanishes outside the marked tiles
    (what the sampler's scatter produces), is bit-identical everywhere, regulariser gradient included."""
    from trinerflet_b200.triplane_encoder import build_planes_with_abs
    R = n0 * 2 ** levels
    T = R // 32
    g = torch.Generator().manual_seed(5)
    flags = torch.rand(3, T, T, generator=g) < density
    if density > 0:
        flags[:, T // 4: T // 2 + 1, T // 3: T // 2 + 1] = True
    plan = _plan_for(flags, R, n0, levels, C)
    pf, coefs = _rand_coefs(C, n0, levels, seed=2)
    with torch.no_grad():
        coefs[-1][:, :, :, ::3, ::4] = 0.0                      # exact zeros: sign(0) = 0 in the regulariser gradient
    mask = flags.cuda().repeat_interleave(32, 1).repeat_interleave(32, 2)[:, None]       # [3,1,R,R]
    gout = torch.randn(3, C, R, R, generator=g).cuda() * mask
    w_abs = torch.rand(levels, generator=g).cuda()
    res = []
    for p in (None, plan):
        pf_g = cl_planes(pf.cuda()).requires_grad_(True)
        coefs_g = [cl_coefs(c.cuda()).requires_grad_(True) for c in coefs]
        out, abs_sums = build_planes_with_abs(pf_g, coefs_g, p)
        ((out * gout).sum() + (abs_sums * w_abs).sum()).backward()
        res.append((out.detach(), abs_sums.detach(), pf_g.grad, [c.grad for c in coefs_g]))
    (o_d, a_d, gp_d, gc_d), (o_s, a_s, gp_s, gc_s) = res
    assert torch.equal(torch.where(mask, o_s, 0.0), torch.where(mask, o_d, 0.0))
    assert rel_l2(a_s, a_d) <= 1e-5
    assert torch.equal(gp_s, gp_d)
    for a, b in zip(gc_s, gc_d):
        assert torch.equal(a, b)
    assert plan.stats["tile_fraction"] == pytest.approx(float(flags.float().mean()))
