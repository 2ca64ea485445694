"""CPU: the oracle restatement (oracle/wavelet.py + oracle/field.py) against the fixtures produced by the
reference's own TriPlaneVolume / NeRFNetwork modules (tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

from oracle import field as of, wavelet as ow


def _t(a):
    return torch.from_numpy(np.asarray(a))


def test_encoder_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "encoder_fp32.npz"))
    pf = _t(g["planes_features"]).requires_grad_(True)
    coefs = [_t(g[f"coef{i}"]).requires_grad_(True) for i in range(2)]
    planes = ow.build_planes(pf, coefs)
    assert torch.equal(planes.detach(), _t(g["planes"]))
    feat = of.sample_planes(planes, _t(g["xyz"]), float(g["bound"]), fp16=False, recip_mul=False)
    assert torch.equal(feat.detach(), _t(g["feat"]))
    (feat * _t(g["wfeat"])).sum().backward()
    assert (pf.grad - _t(g["g_planes_features"])).abs().max().item() < 1e-5
    for i, c in enumerate(coefs):
        assert (c.grad - _t(g[f"g_coef{i}"])).abs().max().item() < 1e-5


def test_field_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "field_fp32.npz"))
    pf = _t(g["planes_features"]).requires_grad_(True)
    coefs = [_t(g[f"coef{i}"]).requires_grad_(True) for i in range(2)]
    Ws = [_t(g[f"W{i}"]).requires_grad_(True) for i in range(1, 6)]
    planes = ow.build_planes(pf, coefs)
    feat = of.sample_planes(planes, _t(g["xyz"]), float(g["bound"]), recip_mul=False)
    sig, rgb, geo = of.mlp_forward(feat, _t(g["dirs"]), Ws, fp16=False)
    assert torch.equal(sig.detach(), _t(g["sigma"])) and torch.equal(rgb.detach(), _t(g["color"]))
    ds, dg = of.density_forward(feat, Ws)
    assert torch.equal(ds.detach(), _t(g["dens_sigma"])) and torch.equal(dg.detach(), _t(g["dens_geo"]))
    ((sig * _t(g["wsig"])).sum() + (rgb * _t(g["wrgb"])).sum()).backward()
    for i, w in enumerate(Ws):
        assert (w.grad - _t(g[f"g_W{i + 1}"])).abs().max().item() < 1e-5
    assert (pf.grad - _t(g["g_planes_features"])).abs().max().item() < 1e-5


def test_fp16_emulation_is_close_to_fp32_and_rounds_outputs():
    g = torch.Generator().manual_seed(0)
    W = of.init_mlp_weights(16, gen=g)
    feat = 0.3 * torch.randn(500, 48, generator=g)
    d = torch.randn(500, 3, generator=g); d = d / d.norm(dim=-1, keepdim=True)
    s32, c32, _ = of.mlp_forward(feat, d, W, fp16=False)
    s16, c16, _ = of.mlp_forward(feat, d, W, fp16=True)
    assert (c16 - c32).abs().max().item() < 5e-3 and ((s16 - s32).abs() / s32).max().item() < 2e-2
    assert torch.equal(c16, c16.half().float())


def test_sh16_orthonormal():
    """16 real SH functions are orthonormal on the sphere (Monte-Carlo quadrature) and [0] = 1/(2 sqrt(pi))."""
    g = torch.Generator().manual_seed(0)
    d = torch.randn(400000, 3, dtype=torch.float64, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    Y = of.sh16(d)
    assert abs(Y[0, 0].item() - 0.28209479177387814) < 1e-15
    G = (Y.T @ Y) * (4 * np.pi / d.shape[0])
    assert (G - torch.eye(16, dtype=torch.float64)).abs().max().item() < 0.02
