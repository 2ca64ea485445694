"""GPU: the opt-in hybrid backward of the 128-wide MLP heads ("large" config, model.wide_fused_backward = True): fused
forward, fused input-gradient chain (tnl_mlp_backward_chain, csrc/mlp.cu in DUMP mode) + library GEMMs for the weight
gradients, against the oracle's fp16-autocast autograd and against the library path the config trains through by default.
The same comparison runs on the host build of the kernels (tests/test_kernels_emu.py, tests/test_host_on_emu.py).

(The file sorts after the other GPU test files on purpose: it was written after this round's GPU budget was spent, its
first run on hardware is the driver's round-end run.)"""
import pytest
import torch

from tests.util import rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("C,M,half", [(32, 20000, False), (48, 5000, True), (16, 777, True)])
def test_field_mlp_wide_heads_hybrid_backward(C, M, half):
    from oracle import field as of
    from trinerflet_b200.network import _FieldMLP
    g = torch.Generator().manual_seed(C)
    W = of.init_mlp_weights(C, 128, 128, gen=g)
    feat = 0.5 * torch.randn(M, 3 * C, generator=g)
    d = torch.randn(M, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    gs, grgb = torch.randn(M, generator=g) * 64.0, torch.randn(M, 3, generator=g) * 64.0
    W_o = [w.clone().requires_grad_(True) for w in W]
    f_o = feat.clone().requires_grad_(True)
    s_o, rgb_o, _ = of.mlp_forward(f_o, d, W_o, fp16=True)
    ((s_o * gs).sum() + (rgb_o * grgb).sum()).backward()
    W_g = [w.clone().cuda().requires_grad_(True) for w in W]
    f_g = (feat.clone().cuda().half() if half else feat.clone().cuda()).requires_grad_(True)
    nv = torch.tensor([M - 5], dtype=torch.int32, device="cuda")
    s_g, rgb_g = _FieldMLP.apply(f_g, d.cuda(), None, *W_g)
    assert (rgb_g.cpu() - rgb_o).abs().max().item() <= 2e-3 and rel_l2(s_g, s_o) <= 2e-3
    ((s_g * gs.cuda()).sum() + (rgb_g * grgb.cuda()).sum()).backward()
    assert f_g.grad.dtype == f_g.dtype and rel_l2(f_g.grad.float(), f_o.grad) <= 1e-2
    for a, b in zip(W_g, W_o):
        assert a.grad.shape == b.grad.shape and rel_l2(a.grad, b.grad) <= 1e-2
    # n_valid: rows past it neither produce outputs nor contribute gradient
    W_h = [w.clone().cuda().requires_grad_(True) for w in W]
    f_h = (feat.clone().cuda().half() if half else feat.clone().cuda()).requires_grad_(True)
    s_h, rgb_h = _FieldMLP.apply(f_h, d.cuda(), nv, *W_h)
    assert float(s_h[M - 5:].abs().sum()) == 0 and torch.equal(s_h[:M - 5], s_g[:M - 5])
    ((s_h * gs.cuda()).sum() + (rgb_h * grgb.cuda()).sum()).backward()
    assert float(f_h.grad[M - 5:].abs().sum()) == 0
    W_p = [w.clone().requires_grad_(True) for w in W]
    s_p, rgb_p, _ = of.mlp_forward(feat[:M - 5], d[:M - 5], W_p, fp16=True)
    ((s_p * gs[:M - 5]).sum() + (rgb_p * grgb[:M - 5]).sum()).backward()
    for a, b in zip(W_h, W_p):
        assert rel_l2(a.grad, b.grad) <= 1e-2


def test_large_config_network_trains_through_the_hybrid_path():
    """NeRFNetwork with hidden 128 under fp16 autocast: wide_fused_backward = True against the default library path"""
    from trinerflet_b200.network import NeRFNetwork
    from trinerflet_b200 import scene
    res = []
    g = torch.Generator().manual_seed(1)
    x = ((torch.rand(30000, 3, generator=g) * 2 - 1) * 1.4).cuda()
    d = torch.randn(30000, 3, generator=g)
    d = (d / d.norm(dim=-1, keepdim=True)).cuda()
    gs, grgb = torch.randn(30000, generator=g).cuda(), torch.randn(30000, 3, generator=g).cuda()
    for wide in (False, True):
        net = NeRFNetwork(bound=1.5, cuda_ray=True, density_thresh=10, min_near=0.2, triplane_channels=16, triplane_resolution=128,
                          triplane_wavelet_levels=2, hidden_dim=128, hidden_dim_color=128).cuda()
        scene.init_model_(net, seed=0)
        net.wide_fused_backward = wide
        net.train()
        with torch.autocast("cuda", dtype=torch.float16):
            sigma, color = net(x, d)
            ((sigma.float() * gs).sum() + (color.float() * grgb).sum()).backward()
        res.append((sigma.detach().float(), color.detach().float(), [p.grad.detach().clone() for p in net.parameters()]))
    (s_a, c_a, g_a), (s_b, c_b, g_b) = res
    assert rel_l2(s_b, s_a) <= 2e-3 and (c_b - c_a).abs().max().item() <= 2e-3
    for a, b in zip(g_a, g_b):
        assert rel_l2(b, a) <= 1e-2
