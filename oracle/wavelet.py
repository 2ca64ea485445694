"""TEST INFRASTRUCTURE ONLY (oracle) -- CPU restatement of the inverse/forward 2-D DWT the
reference obtains from the third-party dependency `pytorch-wavelets==1.3.0` with filter taps
from `PyWavelets==1.4.1` (pinned in /root/reference/requirements2.txt:113,115; neither is
vendored under /root/reference nor installable here: no network).

PARITY UNPINNED for this file: the reference ships no test or golden vector for the wavelet
transform, and the dependency itself cannot be run.  What pins it instead
(tests/test_oracle_wavelet.py): perfect reconstruction against the analysis bank, the shape
contract the reference relies on (triplane_encoder.py:188-206), DC gain, exact adjointness,
filter sums.  The only residual assumption is the sub-band order inside `yh`
(index 0 = high-pass along H), which matters for loading reference checkpoints only.

Reference call sites this follows:
  reconstruction/triplaneencoder/triplane_encoder.py:184  DWTForward(J=1, wave, mode='zero')
  reconstruction/triplaneencoder/triplane_encoder.py:185  DWTInverse(wave, mode='zero')
  reconstruction/triplaneencoder/triplane_encoder.py:394  x = self.idwt((yl, [yh]))
Published algorithm restated (pytorch_wavelets/dwt/lowlevel.py, v1.3.0):
  sfb1d zero mode:  y = conv_transpose2d(lo, g0, stride 2, padding L-2, groups C)
                      + conv_transpose2d(hi, g1, ...)            (filters NOT flipped)
  SFB2D:            lh, hl, hh = unbind(yh, dim=2)
                    lo = sfb1d(ll, lh, dim=H); hi = sfb1d(hl, hh, dim=H); y = sfb1d(lo, hi, dim=W)
  afb1d zero mode:  conv2d with flipped dec filters, stride 2, padding p//2 where
                    p = 2*(outsize-1) - N + L, outsize = (N + L - 1)//2; one extra zero appended if p odd
  AFB2D:            rows (W) first, then columns (H); outputs ll, (lh, hl, hh)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (trinerflet_b200/) never does.
"""
import torch
import torch.nn.functional as F

# PyWavelets 1.4.1 `pywt.Wavelet('bior6.8')` filter bank, length 18 (values as listed in SURVEY.md 8c).
BIOR68_REC_LO = [
    0.0, 0.0, 0.0, 0.014426282505624435, 0.014467504896790148, -0.07872200106262882,
    -0.04036797903033992, 0.41784910915027457, 0.7589077294536541, 0.41784910915027457,
    -0.04036797903033992, -0.07872200106262882, 0.014467504896790148, 0.014426282505624435,
    0.0, 0.0, 0.0, 0.0]
BIOR68_REC_HI = [
    0.0, -0.0019088317364812906, -0.0019142861290887667, 0.016990639867602342,
    0.01193456527972926, -0.04973290349094079, -0.07726317316720414, 0.09405920349573646,
    0.4207962846098268, -0.8259229974584023, 0.4207962846098268, 0.09405920349573646,
    -0.07726317316720414, -0.04973290349094079, 0.01193456527972926, 0.016990639867602342,
    -0.0019142861290887667, -0.0019088317364812906]
BIOR68_DEC_LO = [
    0.0, 0.0019088317364812906, -0.0019142861290887667, -0.016990639867602342,
    0.01193456527972926, 0.04973290349094079, -0.07726317316720414, -0.09405920349573646,
    0.4207962846098268, 0.8259229974584023, 0.4207962846098268, -0.09405920349573646,
    -0.07726317316720414, 0.04973290349094079, 0.01193456527972926, -0.016990639867602342,
    -0.0019142861290887667, 0.0019088317364812906]
BIOR68_DEC_HI = [
    0.0, 0.0, 0.0, 0.014426282505624435, -0.014467504896790148, -0.07872200106262882,
    0.04036797903033992, 0.41784910915027457, -0.7589077294536541, 0.41784910915027457,
    0.04036797903033992, -0.07872200106262882, -0.014467504896790148, 0.014426282505624435,
    0.0, 0.0, 0.0, 0.0]

WAVELETS = {
    "bior6.8": dict(rec_lo=BIOR68_REC_LO, rec_hi=BIOR68_REC_HI, dec_lo=BIOR68_DEC_LO,
                    dec_hi=BIOR68_DEC_HI, pad=4),  # pad: triplane_encoder.py:174-180
}


def _filt(taps, like, dim):
    t = torch.tensor(taps, dtype=like.dtype, device=like.device)
    shape = [1, 1, 1, 1]
    shape[dim] = t.numel()
    return t.reshape(shape)


def sfb1d(lo, hi, g0, g1, dim):
    """1-D synthesis, zero mode (pytorch_wavelets lowlevel.sfb1d).  lo/hi: [B,C,H,W]."""
    C = lo.shape[1]
    L = len(g0)
    w0 = _filt(g0, lo, dim).repeat(C, 1, 1, 1)
    w1 = _filt(g1, lo, dim).repeat(C, 1, 1, 1)
    s = (2, 1) if dim == 2 else (1, 2)
    pad = (L - 2, 0) if dim == 2 else (0, L - 2)
    return (F.conv_transpose2d(lo, w0, stride=s, padding=pad, groups=C)
            + F.conv_transpose2d(hi, w1, stride=s, padding=pad, groups=C))


def sfb2d(ll, yh, wave="bior6.8"):
    """Single-level inverse 2-D DWT, zero mode.  ll [B,C,h,w], yh [B,C,3,h,w] -> [B,C,2h-16,2w-16]."""
    w = WAVELETS[wave]
    g0, g1 = w["rec_lo"], w["rec_hi"]
    lh, hl, hh = torch.unbind(yh, dim=2)
    lo = sfb1d(ll, lh, g0, g1, dim=2)
    hi = sfb1d(hl, hh, g0, g1, dim=2)
    return sfb1d(lo, hi, g0, g1, dim=3)


def afb1d(x, h0, h1, dim):
    """1-D analysis, zero mode (pytorch_wavelets lowlevel.afb1d).  Returns [B,2C,.,.] interleaved lo/hi."""
    C = x.shape[1]
    N = x.shape[dim]
    L = len(h0)
    f0 = _filt(h0[::-1], x, dim)
    f1 = _filt(h1[::-1], x, dim)
    h = torch.cat([f0, f1] * C, dim=0)
    outsize = (N + L - 1) // 2
    p = 2 * (outsize - 1) - N + L
    if p % 2 == 1:
        x = F.pad(x, (0, 0, 0, 1) if dim == 2 else (0, 1, 0, 0))
    pad = (p // 2, 0) if dim == 2 else (0, p // 2)
    s = (2, 1) if dim == 2 else (1, 2)
    return F.conv2d(x, h, padding=pad, stride=s, groups=C)


def afb2d(x, wave="bior6.8"):
    """Single-level forward 2-D DWT, zero mode -> (ll [B,C,h,w], yh [B,C,3,h,w])."""
    w = WAVELETS[wave]
    h0, h1 = w["dec_lo"], w["dec_hi"]
    B, C = x.shape[:2]
    lohi = afb1d(x, h0, h1, dim=3)
    y = afb1d(lohi, h0, h1, dim=2)
    y = y.reshape(B, C, 4, y.shape[-2], y.shape[-1])
    return y[:, :, 0].contiguous(), y[:, :, 1:].contiguous()


class DWTInverse(torch.nn.Module):
    """Stand-in with the call signature the reference uses (triplane_encoder.py:185,394)."""

    def __init__(self, wave="bior6.8", mode="zero"):
        super().__init__()
        assert mode == "zero"
        self.wave = wave

    def forward(self, coeffs):
        yl, yh = coeffs
        ll = yl
        for h in yh[::-1]:
            if ll.shape[-2] > h.shape[-2]:
                ll = ll[..., :-1, :]
            if ll.shape[-1] > h.shape[-1]:
                ll = ll[..., :-1]
            ll = sfb2d(ll, h, self.wave)
        return ll


class DWTForward(torch.nn.Module):
    """Stand-in for triplane_encoder.py:184 (only used there to probe sub-band shapes)."""

    def __init__(self, J=1, wave="bior6.8", mode="zero"):
        super().__init__()
        assert mode == "zero"
        self.J, self.wave = J, wave

    def forward(self, x):
        yh = []
        ll = x
        for _ in range(self.J):
            ll, h = afb2d(ll, self.wave)
            yh.append(h)
        return ll, yh


def _h(t):
    return t.half().float()


def build_planes_fp16_autocast(planes_features, coefs, wave="bior6.8"):
    """What the reference's RENDER path computes: `get_planes()` is first reached inside
    `torch.cuda.amp.autocast` + no_grad (nerf/utils.py:850-856), so every conv_transpose2d of the IDWT is an autocast-fp16 op:
    operands and filter taps rounded to fp16, fp32 accumulation, fp16 result; the sums of two conv results are fp16 + fp16 -> fp16;
    `2 * x` and F.pad keep the dtype.  Emulated here with fp32 arithmetic on fp16-rounded values.  The result is the fp16 plane
    stack the reference caches for the whole evaluation run (triplane_encoder.py:409-416), returned as fp32 holding fp16 values."""
    w = WAVELETS[wave]
    pad = w["pad"]
    g0 = [float(torch.tensor(v, dtype=torch.float32).half()) for v in w["rec_lo"]]
    g1 = [float(torch.tensor(v, dtype=torch.float32).half()) for v in w["rec_hi"]]

    def sfb1d_h(lo, hi, dim):
        C = lo.shape[1]
        L = len(g0)
        w0 = _filt(g0, lo, dim).repeat(C, 1, 1, 1)
        w1 = _filt(g1, lo, dim).repeat(C, 1, 1, 1)
        s = (2, 1) if dim == 2 else (1, 2)
        p = (L - 2, 0) if dim == 2 else (0, L - 2)
        a = _h(F.conv_transpose2d(_h(lo), w0, stride=s, padding=p, groups=C))
        b = _h(F.conv_transpose2d(_h(hi), w1, stride=s, padding=p, groups=C))
        return _h(a + b)

    x = planes_features.float()
    for yh in coefs:
        yl = F.pad(2 * x, (pad, pad, pad, pad))
        yhp = F.pad(yh.float(), (pad, pad, pad, pad))
        lh, hl, hh = torch.unbind(yhp, dim=2)
        lo = sfb1d_h(yl, lh, 2)
        hi = sfb1d_h(hl, hh, 2)
        x = sfb1d_h(lo, hi, 3)
    return x


def build_planes(planes_features, coefs, wave="bior6.8"):
    """Multilevel reconstruction of the three planes; restates
    TriPlaneVolume.build_planes (triplane_encoder.py:364-405) for the configuration every README
    command uses (all levels learnable, no max_res/max_scale cut, wavelet_base_resolution=0):
    per level  yl = 2*x ; zero-pad yl, yh by `pad` on all four sides ; x = IDWT(yl, [yh]).
    planes_features [3,C,n0,n0]; coefs[l] [3,C,3,n0*2^l,n0*2^l]  ->  [3,C,R,R]."""
    pad = WAVELETS[wave]["pad"]
    x = planes_features
    for yh in coefs:
        yl = 2 * x
        yl = F.pad(yl, (pad, pad, pad, pad))
        yh = F.pad(yh, (pad, pad, pad, pad))
        x = sfb2d(yl, yh, wave)
    return x


def build_planes_limited(planes_features, coefs, max_res=-1, max_scale=-1, get_all_resolutions=False, wave="bior6.8"):
    """TriPlaneVolume.build_planes (triplane_encoder.py:364-405) with its optional arguments, as the reference's tooling calls it
    (save_triplane, nerf/utils.py:1649; get_grid_features, triplane_encoder.py:500): the level loop STOPS at the first level
    whose input side has reached max_res or whose accumulated scale has reached max_scale (:380-383; the zero-coefficient
    continuation is commented out there), so the planes come back at that coarser side; get_all_resolutions collects the input
    of every level visited and the final planes (:377-378, :397-398 -- after a stop the last entry appears twice).
    -> (planes, all_res)"""
    pad = WAVELETS[wave]["pad"]
    x, all_res, current_scale = planes_features, [], 1
    for yh in coefs:
        if get_all_resolutions:
            all_res.append(x)
        yl = 2 * x
        if (max_res > 0 and min(x.shape[2:]) >= max_res) or (max_scale > 0 and current_scale >= max_scale):
            break
        yl = F.pad(yl, (pad, pad, pad, pad))
        yh = F.pad(yh, (pad, pad, pad, pad))
        x = sfb2d(yl, yh, wave)
        current_scale *= 2
    if get_all_resolutions:
        all_res.append(x)
    return x, all_res


def idwt_level_closed_form_1d(x, d, g0, g1):
    """Closed form of one padded 1-D synthesis step, pure Python, tiny inputs:
    y[i] = sum_m 2*x[m]*g0[i+8-2m] + d[m]*g1[i+8-2m],  0 <= i+8-2m < 18,  i in [0, 2n)."""
    n = len(x)
    y = [0.0] * (2 * n)
    for i in range(2 * n):
        acc = 0.0
        for m in range(n):
            k = i + 8 - 2 * m
            if 0 <= k < 18:
                acc += 2.0 * x[m] * g0[k] + d[m] * g1[k]
        y[i] = acc
    return y
