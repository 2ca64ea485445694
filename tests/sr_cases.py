"""Shared body of the super_resolution-encoder parity tests: the same checks run on the GPU
(tests/test_gpu_sr_encoder.py) and, over the host build of the kernels, on CPU (tests/test_sr_encoder.py).

Tolerances (fp32 against the reference's fp32 CPU results in tests/golden/sr_encoder_fp32.npz): relative L2 1e-5 for planes,
features and parameter gradients, 1e-4 for the position gradient (a sum of differences of neighbouring texels)."""
import os

import numpy as np
import torch

from tests.golden import make_sr_golden as G
from tests.util import rel_l2

TOL, TOL_X = 1e-5, 1e-4


def golden():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sr_encoder_fp32.npz"))


def _state(z, prefix):
    return {k[len(prefix):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix)}


def _t(z, key, device):
    return torch.from_numpy(z[key]).to(device)


def check_wavelet_golden(device):
    from trinerflet_b200 import sr_encoder
    z = golden()
    x, w_low, w_high = _t(z, "x", device), _t(z, "w_low", device), _t(z, "w_high", device)
    for lo, hi in G.SCALES:
        enc = sr_encoder.TriPlaneVolume(low_res_scale=lo, high_res_scale=hi, **G.WAVELET)
        enc.load_state_dict(_state(z, "wavelet/state/"), strict=True)      # the reference's keys, NCHW-contiguous tensors
        enc = enc.to(device)
        tag = f"wavelet/{lo}_{hi}/"
        enc.enable_cache = False
        enc.set_double_mode(False)
        with torch.no_grad():
            planes = enc.get_planes()
            assert tuple(planes.shape) == tuple(z[tag + "planes_single_shape"])
            assert rel_l2(planes, _t(z, tag + "planes_low", device)) <= TOL
            assert rel_l2(enc(x), _t(z, tag + "feat_single", device)) <= TOL
        f_low, f_high, grads, gx = G.two_render_step(enc, x, w_low, w_high)
        assert rel_l2(f_low, _t(z, tag + "feat_low", device)) <= TOL
        assert rel_l2(f_high, _t(z, tag + "feat_high", device)) <= TOL
        assert rel_l2(gx, _t(z, tag + "grad_x", device)) <= TOL_X
        for name, g in grads.items():
            ref = _t(z, tag + "grad/" + name, device)
            if float(ref.abs().max()) == 0.0:
                assert float(g.abs().max()) == 0.0, name        # a level no reading reaches
            else:
                assert rel_l2(g, ref) <= TOL, (name, lo, hi)
        enc.enable_cache = True
        enc.set_double_mode(True)
        with torch.no_grad():
            for mode in ("low_res", "high_res"):
                enc.set_resolution_mode(mode)
                planes = enc.get_planes()
                key = tag + ("planes_low" if mode == "low_res" else "planes_high")
                assert tuple(planes.shape) == z[key].shape
                assert rel_l2(planes, _t(z, key, device)) <= TOL
            assert enc.get_planes() is planes                   # served from the cache until reset_cahce()
            enc.reset_cahce()
            assert enc.last_used_planes is None


def check_kplanes_golden(device):
    from trinerflet_b200 import sr_encoder
    z = golden()
    x = _t(z, "x", device)
    for mode in ("concatination", "mul"):
        tag = f"kplanes/{mode}/"
        kp = sr_encoder.KPlaneVolume(features_mode=mode, **G.KPLANES)
        kp.load_state_dict(_state(z, tag + "state/"), strict=True)
        kp = kp.to(device)
        assert kp.n_output_dims == z[tag + "feat"].shape[1]
        xg = x.clone().requires_grad_(True)
        f = kp(xg)
        (f * _t(z, tag + "w", device)).sum().backward()
        assert rel_l2(f, _t(z, tag + "feat", device)) <= TOL
        assert rel_l2(xg.grad, _t(z, tag + "grad_x", device)) <= TOL_X
        for n, p in kp.named_parameters():
            assert rel_l2(p.grad, _t(z, tag + "grad/" + n, device)) <= TOL, n
    for name, cls in (("multiscale", sr_encoder.MultiscaleKPlaneVolume), ("multiscale_mul", sr_encoder.MultiscaleKPlaneMulVolume)):
        tag = name + "/"
        ms = cls(features_mode="concatination", **G.MULTISCALE)
        ms.load_state_dict(_state(z, tag + "state/"), strict=True)
        ms = ms.to(device)
        with torch.no_grad():
            f_low = ms(x)
            ms.set_double_mode(True)
            ms.set_resolution_mode('high_res')
            f_high = ms(x)
        assert rel_l2(f_low, _t(z, tag + "feat_low", device)) <= TOL
        assert rel_l2(f_high, _t(z, tag + "feat_high", device)) <= 4 * TOL      # a product of up to nine sampled values
        assert ms.get_wavelet_features() == [] and tuple(ms.get_planes().shape) == (1, 3, 50, 50)


def check_against_oracle(device, C, R, scale, low, high, M, seed=0, exact=True):
    """the two-render step at sizes the golden file does not hold, against oracle/sr_encoder.py
      * in fp32 -- the arithmetic of the reference (same pixel positions bit for bit): TOL / TOL_X;
      * in fp64 -- the exact value: the fp32 rounding of a sample's pixel position (half an ulp of 1 times R texels) moves
        features and scatter weights by about R * 6e-8 of a texel difference, so the bound grows with R
        (exact=False skips this pass: at 1024^2 the fp64 convolutions take half a minute of host time)."""
    from oracle import sr_encoder as osr
    from trinerflet_b200 import sr_encoder
    gen = torch.Generator().manual_seed(seed)
    enc = sr_encoder.TriPlaneVolume(number_of_features=C, plane_resolution=R, inner_multi_res_scale=scale, low_res_scale=low,
                                    high_res_scale=high)
    G.randomise_(enc, gen)
    enc = enc.to(device)
    x = G.test_points(M, gen)
    w_low, w_high = torch.randn(M, 3 * C, generator=gen), torch.randn(M, 3 * C, generator=gen)
    f_low, f_high, grads, gx = G.two_render_step(enc, x.to(device), w_low.to(device), w_high.to(device))
    passes = [(torch.float32, TOL, TOL_X)] + ([(torch.float64, TOL + 6e-8 * R, TOL_X + 6e-8 * R)] if exact else [])
    for dtype, tol, tol_x in passes:
        pf = enc.planes_features.detach().cpu().to(dtype).contiguous().requires_grad_(True)
        coefs = [p.detach().cpu().to(dtype).contiguous().requires_grad_(True) for p in enc.planes_features_wavelet_coefs]
        x_low, x_high = osr.two_readings(pf, coefs, R, low, high, True)
        xd = x.detach().clone().to(dtype).requires_grad_(True)
        o_low, o_high = osr.encode(x_low, xd), osr.encode(x_high, xd)
        ((o_low * w_low.to(dtype)).sum() + (o_high * w_high.to(dtype)).sum()).backward()
        assert tuple(f_low.shape) == tuple(o_low.shape)
        assert rel_l2(f_low, o_low) <= tol and rel_l2(f_high, o_high) <= tol, dtype
        assert rel_l2(gx, xd.grad) <= tol_x, dtype
        assert rel_l2(grads["planes_features"], pf.grad) <= tol, dtype
        for l, c in enumerate(coefs):
            g = grads[f"planes_features_wavelet_coefs.{l}"]
            if c.grad is None:
                assert float(g.abs().max()) == 0.0
            else:
                assert rel_l2(g, c.grad) <= tol, (l, dtype)


def check_position_gradient_properties(device):
    """facts of the position gradient that do not need a reference: zero along an axis whose coordinate is on or beyond the cube
    face, the fp64 derivative in the interior, zeros(0, 3C) for an empty batch, no second derivative"""
    from trinerflet_b200 import sr_encoder
    gen = torch.Generator().manual_seed(3)
    C, R = 8, 16
    enc = sr_encoder.TriPlaneVolume(number_of_features=C, plane_resolution=R, inner_multi_res_scale=1)
    G.randomise_(enc, gen, 1.0)
    enc = enc.to(device)
    x = torch.rand(64, 3, generator=gen) * 0.9 + 0.05
    x[0] = torch.tensor([1.2, 0.4, 0.6])      # beyond +x
    x[1] = torch.tensor([0.3, 0.0, 0.6])      # on the y = 0 face
    x[2] = torch.tensor([0.3, 0.7, -0.5])     # beyond -z
    xg = x.clone().to(device).requires_grad_(True)
    w = torch.randn(64, 3 * C, generator=gen).to(device)
    f = enc(xg)
    g, = torch.autograd.grad((f * w).sum(), xg, create_graph=False)
    assert float(g[0, 0]) == 0.0 and float(g[1, 1]) == 0.0 and float(g[2, 2]) == 0.0
    assert float(g[0, 1:].abs().min()) > 0.0
    # interior points against autograd through F.grid_sample in fp64 (oracle)
    from oracle import sr_encoder as osr
    planes = enc.planes_features.detach().cpu().double().contiguous()
    xd = x[3:].double().requires_grad_(True)
    (osr.encode(planes, xd) * w[3:].cpu().double()).sum().backward()
    assert rel_l2(g[3:], xd.grad) <= TOL_X
    # empty batch: zeros(0, 3C), as triplane_encoder.py:431-434
    e = enc(torch.zeros(0, 3, device=device))
    assert tuple(e.shape) == (0, 3 * C)
    # like F.grid_sample (:262), the position gradient cannot be differentiated again
    xg2 = x.clone().to(device).requires_grad_(True)
    g2, = torch.autograd.grad((enc(xg2) * w).sum(), xg2, create_graph=True)
    try:
        g2.sum().backward()
        differentiable = True
    except RuntimeError:
        differentiable = False
    assert not differentiable


def check_low_resolution_phase_cost(device, R):
    """double mode off: get_planes() stops at the low-resolution level -- only the coarse levels' kernels are launched (one
    launch per level), which is what makes the application's low-resolution phase cheap"""
    from trinerflet_b200 import _lib, sr_encoder
    enc = sr_encoder.TriPlaneVolume(number_of_features=16, plane_resolution=R, inner_multi_res_scale=16, low_res_scale=4).to(device)
    assert enc._level_split() == (2, 4)
    before = _lib.launch_count
    with torch.no_grad():
        planes = enc.get_planes()
    low_launches = _lib.launch_count - before
    enc.set_double_mode(True)
    enc.set_resolution_mode('high_res')
    before = _lib.launch_count
    with torch.no_grad():
        full = enc.get_planes()
    full_launches = _lib.launch_count - before
    assert tuple(planes.shape) == (3, 16, R // 4, R // 4) and tuple(full.shape) == (3, 16, R, R)
    assert low_launches == 2 and full_launches == 4


def _bilinear_any_order(planes, xyz, bound):
    """plain-torch tri-plane bilinear sampling (align_corners, border clamp) written with index arithmetic, so that autograd can
    differentiate it any number of times w.r.t. the planes -- the ground truth for the high-order op"""
    R = planes.shape[-1]
    u = xyz / bound
    out = []
    for p, (a, b) in enumerate(((0, 2), (0, 1), (1, 2))):
        ix = (((u[:, a] + 1) / 2) * (R - 1)).clamp(0, R - 1)
        iy = (((u[:, b] + 1) / 2) * (R - 1)).clamp(0, R - 1)
        x0, y0 = ix.floor().long(), iy.floor().long()
        x1, y1 = (x0 + 1).clamp(max=R - 1), (y0 + 1).clamp(max=R - 1)
        wx1, wy1 = ix - x0, iy - y0
        wx0, wy0 = 1 - wx1, 1 - wy1
        pl = planes[p]                                          # [C, R, R]
        v = (pl[:, y0, x0] * (wx0 * wy0) + pl[:, y0, x1] * (wx1 * wy0) + pl[:, y1, x0] * (wx0 * wy1) + pl[:, y1, x1] * (wx1 * wy1))
        out.append(v.t())                                       # [M, C]
    return torch.cat(out, dim=-1)


def check_high_order_gradients(device):
    """high_order_gradients = True (the capability of the reference's grid_backward.py): a loss on the GRADIENT w.r.t. the planes
    (gradient-penalty shape) back-propagates through the sampling backward; first and second order against plain torch in fp64"""
    from trinerflet_b200 import sr_encoder
    gen = torch.Generator().manual_seed(9)
    C, R, M = 8, 16, 200
    enc = sr_encoder.TriPlaneVolume(number_of_features=C, plane_resolution=R, inner_multi_res_scale=1, input_pts_in_unit_cube=False,
                                    lbound=1.5)
    G.randomise_(enc, gen, 1.0)
    enc = enc.to(device)
    enc.high_order_gradients = True
    x = (torch.rand(M, 3, generator=gen) * 2 - 1) * 1.6          # some outside the cube: border clamp
    w = torch.randn(M, 3 * C, generator=gen)

    def penalty(sample, planes, xs, ws):
        y = sample(planes, xs)
        s = (y ** 2 * ws).sum()
        g, = torch.autograd.grad(s, planes, create_graph=True)
        return s + (g ** 2).sum(), y, g

    planes = enc.planes_features
    loss, y, g = penalty(lambda pl, xs: enc(xs, bound=1.5), planes, x.to(device), w.to(device))
    loss.backward()
    pd = planes.detach().cpu().double().contiguous().requires_grad_(True)
    loss_d, y_d, g_d = penalty(lambda pl, xs: _bilinear_any_order(pl, xs, 1.5), pd, x.double(), w.double())
    loss_d.backward()
    assert rel_l2(y, y_d) <= TOL and rel_l2(g, g_d) <= TOL
    assert abs(float(loss.detach()) - float(loss_d.detach())) <= TOL * abs(float(loss_d.detach()))
    assert rel_l2(planes.grad, pd.grad) <= 10 * TOL              # second order: two more passes of fp32 sampling / scatter
    # third order exists as well (the ops alternate); the default op refuses the second
    enc.zero_grad()
    y = enc(x.to(device), bound=1.5)
    g1, = torch.autograd.grad((y ** 3).sum(), planes, create_graph=True)
    g2, = torch.autograd.grad((g1 ** 2).sum(), planes, create_graph=True)
    (g2 ** 2).sum().backward()
    assert torch.isfinite(planes.grad).all() and float(planes.grad.abs().max()) > 0
