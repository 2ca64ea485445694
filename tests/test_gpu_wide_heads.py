"""GPU: the 128-wide MLP heads of the "large" config (reference README.md:55: --hidden_dim 128 --hidden_dim_color 128) on the
tcgen05 kernels of csrc/mlp_tc128.cu -- fused forward AND fused backward (weight gradients accumulated in TMEM), no library
GEMM anywhere -- against the oracle's fp16-autocast emulation (oracle/field.py; tolerances of SURVEY.md 8c: sigma rel 2e-3,
rgb abs 2e-3, gradients rel-L2 1e-2)."""
import pytest
import torch

from tests.util import rel_l2

pytestmark = pytest.mark.gpu


def _inputs(C, M, seed):
    from oracle import field as of
    g = torch.Generator().manual_seed(seed)
    W = of.init_mlp_weights(C, 128, 128, gen=g)
    feat = (0.5 * torch.randn(M, 3 * C, generator=g)).half().float()      # fp16-representable: both sides round identically
    d = torch.randn(M, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    gs, grgb = torch.randn(M, generator=g) * 64.0, torch.randn(M, 3, generator=g) * 64.0
    return W, feat, d, gs, grgb


@pytest.mark.parametrize("C,M", [(48, 5000), (48, 128 * 148 * 2 + 77), (32, 20000), (16, 777), (48, 1)])
def test_wide_heads_forward_backward_vs_oracle(C, M):
    from oracle import field as of
    from trinerflet_b200.network import _DensityMLP, _FieldMLP
    W, feat, d, gs, grgb = _inputs(C, M, seed=C + M)
    W_o = [w.clone().requires_grad_(True) for w in W]
    f_o = feat.clone().requires_grad_(True)
    s_o, rgb_o, geo_o = of.mlp_forward(f_o, d, W_o, fp16=True)
    ((s_o * gs).sum() + (rgb_o * grgb).sum()).backward()
    W_g = [w.clone().cuda().requires_grad_(True) for w in W]
    f_g = feat.clone().cuda().half().requires_grad_(True)
    s_g, rgb_g = _FieldMLP.apply(f_g, d.cuda(), None, *W_g)
    assert (rgb_g.cpu() - rgb_o).abs().max().item() <= 2e-3
    assert rel_l2(s_g, s_o) <= 2e-3
    s_d, geo_d = _DensityMLP.apply(f_g.detach(), *W_g)
    assert torch.equal(s_d, s_g)
    assert (geo_d.cpu() - geo_o).abs().max().item() <= 2e-3 * max(1.0, geo_o.abs().max().item())
    ((s_g * gs.cuda()).sum() + (rgb_g * grgb.cuda()).sum()).backward()
    assert f_g.grad.dtype == torch.float16 and rel_l2(f_g.grad.float(), f_o.grad) <= 1e-2
    for a, b in zip(W_g, W_o):
        assert a.grad.shape == b.grad.shape and rel_l2(a.grad, b.grad) <= 1e-2, (tuple(a.shape), rel_l2(a.grad, b.grad))
    # the tcgen05 forward agrees with the mma.sync forward (fp32 feature input) to the last bit: same rounding points
    from trinerflet_b200._lib import MlpDims, call, ptr, stream
    from trinerflet_b200.network import pack_mlp_weights
    import ctypes
    dims = MlpDims(3 * C, 128, 128)
    packed = pack_mlp_weights(dims, W_g)
    s_l = torch.empty(M, device="cuda"); rgb_l = torch.empty(M, 3, device="cuda")
    f32 = feat.cuda().contiguous()
    call("tnl_mlp_forward", ctypes.byref(dims), ptr(packed), ptr(f32), 0, ptr(d.cuda().contiguous()), M, None, ptr(s_l), ptr(rgb_l), None, stream())
    assert torch.equal(s_l, s_g.detach()) and torch.equal(rgb_l, rgb_g.detach())


def test_wide_heads_n_valid_and_fp32_feature_input():
    """rows past *n_valid neither produce outputs nor contribute gradient; an fp32 feature tensor is rounded to fp16 on entry
    (the first Linear's own rounding under autocast) and receives an fp32 gradient"""
    from oracle import field as of
    from trinerflet_b200.network import _FieldMLP
    C, M = 48, 3000
    W, feat, d, gs, grgb = _inputs(C, M, seed=11)
    nv_n = M - 133
    nv = torch.tensor([nv_n], dtype=torch.int32, device="cuda")
    W_h = [w.clone().cuda().requires_grad_(True) for w in W]
    f_h = feat.clone().cuda().requires_grad_(True)                      # fp32 in
    s_h, rgb_h = _FieldMLP.apply(f_h, d.cuda(), nv, *W_h)
    assert float(s_h[nv_n:].abs().sum()) == 0 and float(rgb_h[nv_n:].abs().sum()) == 0
    ((s_h * gs.cuda()).sum() + (rgb_h * grgb.cuda()).sum()).backward()
    assert f_h.grad.dtype == torch.float32 and float(f_h.grad[nv_n:].abs().sum()) == 0
    W_p = [w.clone().requires_grad_(True) for w in W]
    f_p = feat[:nv_n].clone().requires_grad_(True)
    s_p, rgb_p, _ = of.mlp_forward(f_p, d[:nv_n], W_p, fp16=True)
    ((s_p * gs[:nv_n]).sum() + (rgb_p * grgb[:nv_n]).sum()).backward()
    assert rel_l2(s_h[:nv_n], s_p) <= 2e-3 and rel_l2(f_h.grad[:nv_n], f_p.grad) <= 1e-2
    for a, b in zip(W_h, W_p):
        assert rel_l2(a.grad, b.grad) <= 1e-2
    # n_valid = 0: nothing runs, gradients are exact zeros
    W_z = [w.clone().cuda().requires_grad_(True) for w in W]
    f_z = feat.clone().cuda().half().requires_grad_(True)
    s_z, rgb_z = _FieldMLP.apply(f_z, d.cuda(), torch.zeros(1, dtype=torch.int32, device="cuda"), *W_z)
    ((s_z * gs.cuda()).sum() + (rgb_z * grgb.cuda()).sum()).backward()
    assert float(s_z.abs().sum()) == 0 and all(float(w.grad.abs().sum()) == 0 for w in W_z)


def test_large_config_network_trains_fused():
    """NeRFNetwork(hidden 128) under fp16 autocast takes the fused kernels for training (no library GEMM: the MLP weights get
    their gradients from tnl_mlp_backward) and matches the reference op sequence run by torch under the same autocast"""
    from trinerflet_b200 import scene
    from trinerflet_b200.network import NeRFNetwork
    g = torch.Generator().manual_seed(1)
    x = ((torch.rand(30000, 3, generator=g) * 2 - 1) * 1.4).cuda()
    d = torch.randn(30000, 3, generator=g)
    d = (d / d.norm(dim=-1, keepdim=True)).cuda()
    gs, grgb = (torch.randn(30000, generator=g) * 8).cuda(), (torch.randn(30000, 3, generator=g) * 8).cuda()
    res = []
    for fused in (True, False):
        net = NeRFNetwork(bound=1.5, cuda_ray=True, triplane_channels=48, triplane_resolution=256, triplane_wavelet_levels=4,
                          hidden_dim=128, hidden_dim_color=128).cuda()
        scene.init_model_(net, seed=0)
        with torch.autocast("cuda", dtype=torch.float16):
            if fused:
                sigma, rgb = net(x, d)
            else:   # the reference's literal op sequence (network.py:125-147) through torch's autocast GEMMs
                import torch.nn.functional as F
                from trinerflet_b200.activation import trunc_exp
                feat = net.encoder(x, bound=1.5)
                h = F.relu(net.sigma_net[0](feat)); h = net.sigma_net[1](h)
                sigma = trunc_exp(h[..., 0])
                h = torch.cat([net.encoder_dir(d), h[..., 1:]], dim=-1)
                h = F.relu(net.color_net[0](h)); h = F.relu(net.color_net[1](h))
                rgb = torch.sigmoid(net.color_net[2](h))
        ((sigma.float() * gs).sum() + (rgb.float() * grgb).sum()).backward()
        res.append((sigma.detach().float(), rgb.detach().float(), [p.grad.detach().clone() for p in net.parameters()]))
    (s_a, c_a, g_a), (s_b, c_b, g_b) = res
    assert rel_l2(s_a, s_b) <= 2e-3 and (c_a - c_b).abs().max().item() <= 2e-3
    for a, b in zip(g_a, g_b):
        assert rel_l2(a, b) <= 1e-2
