"""GPU parity: fused sigma/color MLP kernels and the whole field (encoder + MLP) against the oracle.
Stated tolerances: sigma/rgb under fp16 autocast <= 2e-3 (abs for rgb in [0,1], relative for sigma),
gradients rel-L2 <= 1e-2 (fp16); fp32 path 1e-5 / 1e-4."""
import os

import numpy as np
import pytest
import torch

from tests.util import rel_l2, rel_linf

pytestmark = pytest.mark.gpu
TOL_RGB = 2e-3
TOL_SIGMA_REL = 2e-3
TOL_GRAD = 1e-2


def _mlp_inputs(C, M, seed=0, hidden=64):
    from oracle import field as of
    g = torch.Generator().manual_seed(seed)
    W = of.init_mlp_weights(C, hidden, hidden, gen=g)
    feat = 0.5 * torch.randn(M, 3 * C, generator=g)
    d = torch.randn(M, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    return W, feat, d


@pytest.mark.parametrize("C,hidden,M", [(16, 64, 4099), (32, 64, 10000), (48, 64, 777), (32, 128, 3000), (48, 128, 1000)])
def test_mlp_forward_fp16(C, hidden, M):
    from oracle import field as of
    from trinerflet_b200.network import _DensityMLP, _FieldMLP
    W, feat, d = _mlp_inputs(C, M, hidden=hidden)
    s_o, rgb_o, geo_o = of.mlp_forward(feat, d, W, fp16=True)
    Wg = [w.cuda() for w in W]
    s_g, rgb_g = _FieldMLP.apply(feat.cuda(), d.cuda(), None, *Wg)
    assert (rgb_g.cpu() - rgb_o).abs().max().item() <= TOL_RGB
    assert ((s_g.cpu() - s_o).abs() / s_o.abs().clamp_min(1e-3)).max().item() <= 2 * TOL_SIGMA_REL
    assert rel_l2(s_g, s_o) <= TOL_SIGMA_REL
    s_d, geo_d = _DensityMLP.apply(feat.cuda(), *Wg)
    assert torch.equal(s_d, s_g)
    s_h, rgb_h = _FieldMLP.apply(feat.cuda().half(), d.cuda(), None, *Wg)     # fp16 feature stream: same rounding point
    assert torch.equal(s_h, s_g) and torch.equal(rgb_h, rgb_g)
    assert (geo_d.cpu() - geo_o).abs().max().item() <= TOL_RGB * max(1.0, geo_o.abs().max().item())


@pytest.mark.parametrize("C,M", [(16, 4099), (32, 20000), (48, 1500)])
def test_mlp_backward_fp16(C, M):
    from oracle import field as of
    from trinerflet_b200.network import _FieldMLP
    W, feat, d = _mlp_inputs(C, M, seed=3)
    g = torch.Generator().manual_seed(9)
    gs = torch.randn(M, generator=g) * 64.0          # loss-scaled gradients, as GradScaler produces
    grgb = torch.randn(M, 3, generator=g) * 64.0
    W_o = [w.clone().requires_grad_(True) for w in W]
    f_o = feat.clone().requires_grad_(True)
    s_o, rgb_o, _ = of.mlp_forward(f_o, d, W_o, fp16=True)
    ((s_o * gs).sum() + (rgb_o * grgb).sum()).backward()
    W_g = [w.cuda().requires_grad_(True) for w in W]
    f_g = feat.cuda().requires_grad_(True)
    s_g, rgb_g = _FieldMLP.apply(f_g, d.cuda(), None, *W_g)
    ((s_g * gs.cuda()).sum() + (rgb_g * grgb.cuda()).sum()).backward()
    assert rel_l2(f_g.grad, f_o.grad) <= TOL_GRAD
    for a, b in zip(W_g, W_o):
        assert rel_l2(a.grad, b.grad) <= TOL_GRAD, (a.shape, rel_l2(a.grad, b.grad))
    # fp16 feature stream in, fp16 feature gradient out: identical rounding points, identical results
    W_f = [w.cuda().requires_grad_(True) for w in W]
    f_f = feat.cuda().half().requires_grad_(True)
    s_f, rgb_f = _FieldMLP.apply(f_f, d.cuda(), None, *W_f)
    ((s_f * gs.cuda()).sum() + (rgb_f * grgb.cuda()).sum()).backward()
    assert f_f.grad.dtype == torch.float16 and torch.equal(f_f.grad.float(), f_g.grad)
    for a, b in zip(W_f, W_g):
        assert rel_l2(a.grad, b.grad) <= 1e-5
    # n_valid: only the first rows contribute
    nv = torch.tensor([M // 3], dtype=torch.int32, device="cuda")
    W_h = [w.cuda().requires_grad_(True) for w in W]
    f_h = feat.cuda().requires_grad_(True)
    s_h, rgb_h = _FieldMLP.apply(f_h, d.cuda(), nv, *W_h)
    assert float(s_h[M // 3:].abs().sum()) == 0 and torch.equal(s_h[:M // 3], s_g[:M // 3])
    ((s_h * gs.cuda()).sum() + (rgb_h * grgb.cuda()).sum()).backward()
    W_p = [w.clone().requires_grad_(True) for w in W]
    s_p, rgb_p, _ = of.mlp_forward(feat[:M // 3], d[:M // 3], W_p, fp16=True)
    ((s_p * gs[:M // 3]).sum() + (rgb_p * grgb[:M // 3]).sum()).backward()
    for a, b in zip(W_h, W_p):
        assert rel_l2(a.grad, b.grad) <= TOL_GRAD
    # the same through the fp16 feature stream (tcgen05 kernels): rows past n_valid neither produce nor receive anything
    W_q = [w.cuda().requires_grad_(True) for w in W]
    f_q = feat.cuda().half().requires_grad_(True)
    s_q, rgb_q = _FieldMLP.apply(f_q, d.cuda(), nv, *W_q)
    assert torch.equal(s_q, s_h) and torch.equal(rgb_q, rgb_h)
    ((s_q * gs.cuda()).sum() + (rgb_q * grgb.cuda()).sum()).backward()
    assert torch.equal(f_q.grad[:M // 3].float(), f_h.grad[:M // 3])
    for a, b in zip(W_q, W_h):
        assert rel_l2(a.grad, b.grad) <= 1e-5


def test_oracle_fp16_emulation_matches_cuda_autocast():
    """Pins the oracle's explicit fp16 rounding points against the same graph under real torch.autocast('cuda')."""
    from oracle import field as of
    W, feat, d = _mlp_inputs(32, 5000, seed=4)
    s_o, rgb_o, _ = of.mlp_forward(feat, d, W, fp16=True)
    Wc = [w.cuda() for w in W]
    with torch.autocast("cuda", dtype=torch.float16):
        s_c, rgb_c, _ = of.mlp_forward(feat.cuda(), d.cuda(), Wc, fp16=False)
    assert rgb_c.dtype == torch.float16
    assert (rgb_c.float().cpu() - rgb_o).abs().max().item() <= 1.5e-3   # one fp16 ulp of values in [0.5, 1)
    assert rel_l2(s_c.float(), s_o) <= 2e-3


def test_network_fp32_golden(golden_dir):
    """NeRFNetwork on the GPU without autocast vs the reference's own modules (fixture generated on CPU, fp32).
    C = 4 in the fixture -> zero-padded to 16 channels (the extra channels have zero coefficients and zero W1 columns)."""
    from trinerflet_b200.network import NeRFNetwork
    g = np.load(os.path.join(golden_dir, "field_fp32.npz"))
    C, R, S, bound = int(g["C"]), int(g["R"]), int(g["S"]), float(g["bound"])
    Cp = 16
    net = NeRFNetwork(bound=bound, cuda_ray=True, density_thresh=10, triplane_channels=Cp, triplane_resolution=R,
                      triplane_wavelet_levels=S).cuda()
    assert set(str(k) for k in g["state_dict_keys"]) <= set(net.state_dict().keys())
    with torch.no_grad():
        net.encoder.planes_features.zero_()
        net.encoder.planes_features[:, :C].copy_(torch.from_numpy(g["planes_features"]))
        for i, p in enumerate(net.encoder.planes_features_wavelet_coefs):
            p.zero_(); p[:, :C].copy_(torch.from_numpy(g[f"coef{i}"]))
        W1 = torch.zeros(64, 3 * Cp)
        W1.view(64, 3, Cp)[:, :, :C] = torch.from_numpy(g["W1"]).view(64, 3, C)
        net.sigma_net[0].weight.copy_(W1)
        net.sigma_net[1].weight.copy_(torch.from_numpy(g["W2"]))
        for i in range(3):
            net.color_net[i].weight.copy_(torch.from_numpy(g[f"W{i + 3}"]))
    xyz, dirs = torch.from_numpy(g["xyz"]).cuda(), torch.from_numpy(g["dirs"]).cuda()
    sigma, color = net(xyz, dirs)
    assert rel_linf(sigma, torch.from_numpy(g["sigma"])) <= 1e-4
    assert (color.cpu() - torch.from_numpy(g["color"])).abs().max().item() <= 1e-5
    dens = net.density(xyz)
    assert rel_linf(dens["sigma"], torch.from_numpy(g["dens_sigma"])) <= 1e-4
    ((sigma * torch.from_numpy(g["wsig"]).cuda()).sum() + (color * torch.from_numpy(g["wrgb"]).cuda()).sum()).backward()
    assert rel_l2(net.sigma_net[1].weight.grad, torch.from_numpy(g["g_W2"])) <= 1e-3
    assert rel_l2(net.color_net[1].weight.grad, torch.from_numpy(g["g_W4"])) <= 1e-3
    assert rel_l2(net.encoder.planes_features.grad[:, :C], torch.from_numpy(g["g_planes_features"])) <= 1e-3
    for i, p in enumerate(net.encoder.planes_features_wavelet_coefs):
        assert rel_l2(p.grad[:, :C], torch.from_numpy(g[f"g_coef{i}"])) <= 1e-3


def test_field_fp16_end_to_end():
    """encoder + fused MLP under autocast vs the oracle's fp16 emulation (C=32, R=256, 3 levels)."""
    from oracle import field as of, wavelet as ow
    from trinerflet_b200 import scene
    from trinerflet_b200.network import NeRFNetwork
    C, R, S, bound, M = 32, 256, 8, 1.5, 30000
    net = NeRFNetwork(bound=bound, cuda_ray=True, density_thresh=10, triplane_channels=C, triplane_resolution=R,
                      triplane_wavelet_levels=S).cuda()
    scene.init_model_(net, seed=1)
    g = torch.Generator().manual_seed(2)
    xyz = (torch.rand(M, 3, generator=g) * 2 - 1) * bound
    d = torch.randn(M, 3, generator=g); d = d / d.norm(dim=-1, keepdim=True)
    gs, grgb = torch.randn(M, generator=g) * 16, torch.randn(M, 3, generator=g) * 16
    with torch.autocast("cuda", dtype=torch.float16):
        sigma, color = net(xyz.cuda(), d.cuda())
    ((sigma * gs.cuda()).sum() + (color.float() * grgb.cuda()).sum()).backward()
    pf = net.encoder.planes_features.detach().cpu().contiguous().requires_grad_(True)
    coefs = [p.detach().cpu().contiguous().requires_grad_(True) for p in net.encoder.planes_features_wavelet_coefs]
    W = [w.detach().cpu().clone().requires_grad_(True) for w in net._weights()]
    planes = ow.build_planes(pf, coefs)
    s_o, rgb_o = of.field_forward(planes, xyz, d, W, bound, fp16=True, recip_mul=True)
    ((s_o * gs).sum() + (rgb_o * grgb).sum()).backward()
    assert (color.float().cpu() - rgb_o).abs().max().item() <= TOL_RGB
    assert rel_l2(sigma, s_o) <= TOL_SIGMA_REL
    for a, b in zip(net._weights(), W):
        assert rel_l2(a.grad, b.grad) <= TOL_GRAD
    assert rel_l2(net.encoder.planes_features.grad, pf.grad) <= TOL_GRAD
    for a, b in zip(net.encoder.planes_features_wavelet_coefs, coefs):
        assert rel_l2(a.grad, b.grad) <= TOL_GRAD
