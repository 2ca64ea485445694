"""CPU: pins the wavelet oracle (oracle/wavelet.py) with the known-answer tests of SURVEY.md 8c -- the reference has
no test of its own for this path and pytorch_wavelets cannot be installed, so these are what anchors it."""
import numpy as np
import torch

from oracle import wavelet as W


def test_filter_sums():
    s2 = np.sqrt(2.0)
    assert abs(sum(W.BIOR68_REC_LO) - s2) < 1e-12 and abs(sum(W.BIOR68_DEC_LO) - s2) < 1e-12
    assert abs(sum(W.BIOR68_REC_HI)) < 1e-12 and abs(sum(W.BIOR68_DEC_HI)) < 1e-12
    g0 = np.array(W.BIOR68_REC_LO)
    assert abs(g0[0::2].sum() - s2 / 2) < 1e-12 and abs(g0[1::2].sum() - s2 / 2) < 1e-12
    assert np.count_nonzero(g0) == 11 and np.count_nonzero(W.BIOR68_REC_HI) == 17


def test_perfect_reconstruction_and_shapes():
    x = torch.randn(2, 3, 64, 48, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    ll, yh = W.afb2d(x)
    assert ll.shape == (2, 3, (64 + 17) // 2, (48 + 17) // 2) and yh.shape == (2, 3, 3, 40, 32)
    y = W.sfb2d(ll, yh)
    assert y.shape == x.shape and (y - x).abs().max().item() < 1e-11


def test_shape_contract_of_init_plane_features():
    """triplane_encoder.py:188-206: DWTForward on ones, crop `pad`, gives the coefficient shapes n, 2n, 4n ..."""
    xfm = W.DWTForward(J=1)
    yl = torch.ones(3, 2, 512, 512)
    shapes = []
    for _ in range(3):
        yl, yh = xfm(yl)
        yl = yl[..., 4:-4, 4:-4]
        shapes.append(tuple(yh[0][..., 4:-4, 4:-4].shape))
    assert shapes == [(3, 2, 3, 256, 256), (3, 2, 3, 128, 128), (3, 2, 3, 64, 64)] and yl.shape[-1] == 64


def test_dc_gain_and_border():
    c = 0.7
    pf = torch.full((3, 1, 32, 32), c, dtype=torch.float64)
    p = W.build_planes(pf, [torch.zeros(3, 1, 3, 32, 32, dtype=torch.float64)])
    assert p.shape == (3, 1, 64, 64)
    assert (p[..., 8:-8, 8:-8] - c).abs().max().item() < 1e-13
    assert (p - c).abs().max().item() <= 0.76 * c                # zero padding disturbs only an 8-px border (corner: 1 - 0.5^2)


def test_closed_form_and_adjoint():
    g = torch.Generator().manual_seed(1)
    n = 24
    x, d = torch.randn(n, dtype=torch.float64, generator=g), torch.randn(n, dtype=torch.float64, generator=g)
    pad = torch.nn.functional.pad
    ref = W.sfb1d(pad((2 * x).view(1, 1, n, 1), (0, 0, 4, 4)), pad(d.view(1, 1, n, 1), (0, 0, 4, 4)),
                  W.BIOR68_REC_LO, W.BIOR68_REC_HI, 2).flatten()
    cf = torch.tensor(W.idwt_level_closed_form_1d(x.tolist(), d.tolist(), W.BIOR68_REC_LO, W.BIOR68_REC_HI), dtype=torch.float64)
    assert (ref - cf).abs().max().item() < 1e-14
    pf = torch.randn(3, 2, 16, 16, dtype=torch.float64, generator=g, requires_grad=True)
    c0 = torch.randn(3, 2, 3, 16, 16, dtype=torch.float64, generator=g, requires_grad=True)
    c1 = torch.randn(3, 2, 3, 32, 32, dtype=torch.float64, generator=g, requires_grad=True)
    y = W.build_planes(pf, [c0, c1])
    gy = torch.randn(y.shape, dtype=torch.float64, generator=g)
    y.backward(gy)
    lhs = (y.detach() * gy).sum()
    rhs = (pf.detach() * pf.grad).sum() + (c0.detach() * c0.grad).sum() + (c1.detach() * c1.grad).sum()
    assert abs(lhs - rhs) < 1e-10 * abs(lhs)


def test_bior68_taps_are_the_6_8_biorthogonal_pair():
    """Known-answer anchor for the embedded filter constants (no PyWavelets here to compare against): 'bior6.8' means the
    reconstruction low-pass has 6 zeros at z = -1 and the decomposition low-pass 8, i.e. the dual high-pass filters have
    exactly 6 resp. 8 vanishing moments.  Together with perfect reconstruction (tested above) the product filter
    rec_lo * dec_lo is then the unique order-14 max-flat half-band filter of length 27 -- a typo in any tap breaks this."""
    import numpy as np
    from oracle import wavelet as ow
    k = np.arange(18, dtype=np.float64) - 8.5
    rec_hi, dec_hi = np.array(ow.BIOR68_REC_HI, dtype=np.float64), np.array(ow.BIOR68_DEC_HI, dtype=np.float64)
    rec_lo, dec_lo = np.array(ow.BIOR68_REC_LO, dtype=np.float64), np.array(ow.BIOR68_DEC_LO, dtype=np.float64)
    for f, order in ((rec_hi, 8), (dec_hi, 6)):
        scale = np.abs(f).sum()
        for p in range(order):
            assert abs((k ** p * f).sum()) <= 1e-8 * scale * 8.5 ** p
        assert abs((k ** order * f).sum()) > 1.0            # ... and no more than that
    # half-band property of the product filter: every second tap vanishes except the centre one (= 1)
    prod = np.convolve(rec_lo, dec_lo)
    centre = int(np.argmax(np.abs(prod)))
    assert abs(prod[centre] - 1.0) <= 1e-12
    others = prod[(np.arange(prod.size) - centre) % 2 == 0]
    assert np.abs(others).sum() - abs(prod[centre]) <= 1e-12
    assert np.count_nonzero(np.abs(prod) > 1e-15) <= 27
    # symmetric filters (linear phase)
    nz = lambda f: f[np.flatnonzero(f)[0]: np.flatnonzero(f)[-1] + 1]
    assert np.allclose(nz(rec_lo), nz(rec_lo)[::-1]) and np.allclose(nz(dec_lo), nz(dec_lo)[::-1])
    assert nz(rec_lo).size == 11 and nz(dec_lo).size == 17
