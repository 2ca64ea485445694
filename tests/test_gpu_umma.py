"""tcgen05 descriptor conventions (csrc/umma.cuh): single products through the probe entry point vs numpy.
Integer-valued fp16 operands make every product and partial sum exact, so the comparison is bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _probe(M, N, K, a_mn, b_mn, ncols):
    from trinerflet_b200._lib import call, ptr, stream
    g = torch.Generator().manual_seed(M + 3 * N + 7 * K)
    A = (torch.randint(-4, 5, (M, K), generator=g).float() / 4).half()
    B = (torch.randint(-4, 5, (N, K), generator=g).float() / 4).half()
    D = (A.float() @ B.float().t()).numpy()
    As = (A.t().contiguous() if a_mn else A.contiguous()).cuda()
    Bs = (B.t().contiguous() if b_mn else B.contiguous()).cuda()
    out = torch.zeros(128, ncols, device="cuda")
    call("tnl_umma_probe", ptr(As), As.shape[0], As.shape[1], ptr(Bs), Bs.shape[0], Bs.shape[1], int(a_mn), int(b_mn), M, N, K,
         ptr(out), ncols, stream())
    torch.cuda.synchronize()
    return D, out.cpu().numpy()


@pytest.mark.parametrize("N,K,b_mn", [(64, 96, False), (16, 64, False), (64, 16, True), (96, 64, True), (48, 64, True)])
def test_m128_products(N, K, b_mn):
    D, out = _probe(128, N, K, False, b_mn, N)
    assert np.array_equal(out[:, :N], D)        # accumulator row i lives in TMEM lane i


@pytest.mark.parametrize("N", [96, 64, 32, 16])
def test_m64_weight_gradient_products(N):
    D, out = _probe(64, N, 128, True, True, N)
    lanes = np.array([(i // 16) * 32 + i % 16 for i in range(64)])
    assert np.array_equal(out[lanes, :N], D)    # rows 16w..16w+15 live in lanes 32w..32w+15


@pytest.mark.parametrize("N,K,a_mn,b_mn", [(144, 128, True, True), (128, 128, True, True), (32, 128, True, True), (16, 128, True, True),
                                            (144, 128, False, True), (128, 16, False, True), (16, 128, False, True), (128, 144, False, False)])
def test_m128_products_of_the_wide_heads(N, K, a_mn, b_mn):
    """the product shapes of csrc/mlp_tc128.cu: M = 128 weight-gradient products (both operands MN-major, contraction over the
    128 points), N = 144 (in_dim of C = 48), K = 144"""
    D, out = _probe(128, N, K, a_mn, b_mn, N if N % 8 == 0 else N + 8 - N % 8)
    assert np.array_equal(out[:, :N], D)
