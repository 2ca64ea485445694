"""TEST INFRASTRUCTURE ONLY (oracle) -- CPU restatement (torch fp32 / fp64 on CPU tensors) of the encoders of the reference's
super_resolution application, /root/reference/super_resolution/threestudio/models/triplaneencoder/triplane_encoder.py:

  two_readings      TriPlaneVolume.get_planes :268-340 -- the level loop with its two stops
  encode            TriPlaneVolume.sample_from_planes / forward :347-369, :421-436 -- unit-cube mapping + F.grid_sample
  kplanes           KPlaneVolume.forward :486-497
  multiscale        MultiscaleKPlaneVolume.forward :523-527
  multiscale_mul    MultiscaleKPlaneMulVolume.forward :569-578

Pinned against the reference's own classes (imported from /root/reference by tests/golden/make_sr_golden.py, which writes
tests/golden/sr_encoder_fp32.npz; tests/test_sr_encoder.py repeats the comparison live when /root/reference is present).
The IDWT arithmetic is oracle/wavelet.py's (see its header for what pins it).  Never imported by the product package."""
import torch

from . import field, wavelet


def two_readings(planes_features, coefs, plane_resolution, low_res_scale, high_res_scale, double_mode, wave="bior6.8"):
    """-> (x_low_res, x_high_res or None), following the loop statement by statement (:283-333, no activation,
    wavelet_base_resolution = 0)."""
    pad = wavelet.WAVELETS[wave]["pad"]
    F = torch.nn.functional
    low_res = plane_resolution / low_res_scale
    high_res = plane_resolution / high_res_scale
    x_low = x_high = None
    x = planes_features
    for yh in coefs:
        if min(x.shape[2:]) >= low_res and x_low is None:
            x_low = x
            if not double_mode:
                break
        if min(x.shape[2:]) >= high_res and x_high is None:
            x_high = x
            break
        yl = F.pad(2 * x, (pad, pad, pad, pad))
        x = wavelet.sfb2d(yl, F.pad(yh, (pad, pad, pad, pad)), wave)
    if x_low is None:
        x_low = x
    if x_high is None and double_mode:
        x_high = x
    return x_low, x_high


def encode(planes, coordinates, lbound=1.0, unit_cube=True):
    """[M,3] -> [M, 3C]; coordinates in the unit cube are mapped to [-lbound, lbound] (:364-365), then divided by lbound again
    inside sample_from_planes_aux (:255) -- true division, as on the CPU the goldens were made on."""
    if coordinates.shape[0] == 0:
        return torch.zeros(0, 3 * planes.shape[1], dtype=coordinates.dtype)
    if unit_cube:
        coordinates = (coordinates * 2 - 1) * lbound
    return field.sample_planes(planes, coordinates, lbound, fp16=False, recip_mul=False)


def kplanes(planes_per_level, coordinates, features_mode, lbound=1.0):
    out = []
    for planes in planes_per_level:
        f = encode(planes, coordinates, lbound)
        if features_mode == "mul":
            f = f.view(f.shape[0], 3, planes.shape[1])
            f = f[:, 0] * f[:, 1] * f[:, 2]
        out.append(f)
    return torch.cat(out, dim=-1)


def multiscale(low_planes, high_planes, coordinates, features_mode, high):
    res = kplanes(low_planes, coordinates, features_mode)
    if high:
        res = torch.cat([res, kplanes(high_planes, coordinates, features_mode)], dim=-1)
    return res


def _mul_all(x, channels):
    x = x.view(x.shape[0], -1, 3 * channels)
    res = x[:, 0]
    for i in range(1, x.shape[1]):
        res = res * x[:, i]
    return res


def multiscale_mul(low_planes, high_planes, coordinates, features_mode, high):
    """Note the reference reshapes to rows of 3*channels, so with features_mode='mul' (level features of `channels` values)
    the product runs over groups of three LEVELS; restated as written."""
    channels = low_planes[0].shape[1]
    res = _mul_all(kplanes(low_planes, coordinates, features_mode), channels)
    if high:
        res = res * _mul_all(kplanes(high_planes, coordinates, features_mode), channels)
    return res
