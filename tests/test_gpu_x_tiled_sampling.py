"""GPU: the opt-in tile-binned sampling kernels (csrc/tsample.cu: tnl_tap_sort / tnl_tsample_forward / tnl_tsample_backward)
against the default point-ordered kernels (csrc/sample.cu): forward bit-identical, backward equal up to the order of the
float additions; and a whole training step (fp16 autocast, fused MLP kernels, work-list IDWT) with
encoder.tiled_sampling = True against the default step.  The same comparisons run on the host build of the kernels in
tests/test_kernels_emu.py and tests/test_host_on_emu.py.

(The file sorts after the other GPU test files on purpose: it was written after this round's GPU budget was spent, its
first run on hardware is the driver's round-end run.)"""
import numpy as np
import pytest
import torch

from tests.util import cl_planes, rel_l2

pytestmark = pytest.mark.gpu
BOUND = 1.5


@pytest.mark.parametrize("C,R,fp16,use_list", [(16, 64, False, False), (32, 256, True, True), (48, 128, True, False), (32, 2048, True, True)])
def test_tiled_kernels_equal_point_ordered_kernels(C, R, fp16, use_list):
    from trinerflet_b200 import _lib
    from trinerflet_b200._lib import call, ptr, stream
    from trinerflet_b200.triplane_encoder import _inv_bound, cl_empty_planes, tap_sort
    g = torch.Generator().manual_seed(C + R)
    M = 200_000 if R == 2048 else 20_000
    planes = cl_empty_planes(C, R, device="cuda").normal_()          # [3,C,R,R] logical, channels-last storage
    xyz = ((torch.rand(M, 3, generator=g) * 2 - 1) * BOUND * (0.5 if R == 2048 else 1.02)).cuda()
    nv = torch.tensor([M - 77], dtype=torch.int32, device="cuda")
    inv = _inv_bound(BOUND)
    perm, bins = tap_sort(xyz, BOUND, R, fp16, nv)
    assert torch.equal(torch.sort(perm.long()).values.cpu(), torch.arange(M))
    G = R // 32
    ids = cnt = None
    cap = 0
    if use_list:
        ids = torch.cat([torch.randperm(3 * G * G, generator=g), torch.full((5,), 10 ** 6)]).int().cuda()
        cnt = torch.tensor([3 * G * G], dtype=torch.int32, device="cuda")
        cap = ids.numel()
    pl = planes.permute(0, 2, 3, 1)                                   # dense [3][R][R][C] view
    assert pl.is_contiguous()
    for half in (0, 1):
        dt = torch.float16 if half else torch.float32
        ref = torch.full((M, 3 * C), float("nan"), device="cuda", dtype=dt)
        call("tnl_sample_planes_forward", ptr(pl), ptr(xyz), M, R, C, inv, int(fp16), ptr(nv), None, ptr(ref), half, stream())
        feat = torch.full((M, 3 * C), float("nan"), device="cuda", dtype=dt)
        call("tnl_tsample_forward", ptr(pl), ptr(xyz), M, R, C, inv, int(fp16), ptr(nv), ptr(perm), ptr(bins), ptr(ids), ptr(cnt), cap,
             ptr(feat), half, stream())
        assert torch.equal(feat, ref)
        gf = torch.randn(M, 3 * C, generator=g).cuda().to(dt)
        gp_ref = torch.zeros(3, R, R, C, device="cuda")
        call("tnl_sample_planes_backward", ptr(gf), half, ptr(xyz), M, R, C, inv, int(fp16), ptr(nv), None, ptr(gp_ref), stream())
        gp = torch.full((3, R, R, C), float("nan"), device="cuda")   # no zero fill: every tile is written
        halo = torch.full((_lib.load().tnl_tsample_backward_workspace(R, C) // 4,), float("nan"), device="cuda")
        tmap = None
        if use_list:
            tmap = torch.zeros(3 * G * G, dtype=torch.uint8, device="cuda")
            tmap[ids[:3 * G * G].long()] = 1
        call("tnl_tsample_backward", ptr(gf), half, ptr(xyz), M, R, C, inv, int(fp16), ptr(perm), ptr(bins), ptr(ids), ptr(cnt), cap,
             ptr(tmap), ptr(gp), ptr(halo), halo.numel() * 4, stream())
        assert bool(torch.isfinite(gp).all())
        assert (gp - gp_ref).abs().max().item() <= 3e-5 * gp_ref.abs().max().item()


def test_training_step_with_tiled_sampling_equals_default_step():
    from tests.test_gpu_train import _model
    from trinerflet_b200 import scene, trainer
    sc = scene.make_scene()
    b = [tuple(t.cuda() for t in scene.sample_batch(sc, 4096, torch.Generator().manual_seed(11 + i))) for i in range(2)]
    res = []
    for tiled in (False, True):
        net = _model("cpu")                      # C=16, R=512, 3 levels: the work-list path is active
        net.encoder.tiled_sampling = tiled
        ts = trainer.TrainStep(net, trainer.default_opt(), None, world_size=1)
        ts.plan_on_any_device = True             # (no effect on CUDA tensors; lets the CPU dry run of this file take the same path)
        torch.manual_seed(0)
        ts.forward_backward(*b[0], update_grid=False)
        net.mean_count = int(net.step_counter[0, 0].item())
        net.local_step = 0
        net.zero_grad(set_to_none=True)
        torch.manual_seed(5)
        loss = ts.forward_backward(*b[1], update_grid=False)
        res.append((float(loss), [p.grad.detach().clone() for p in net.parameters()], net))
    (l_a, g_a, _), (l_b, g_b, net) = res
    assert net.encoder.sampling_tiles is None            # valid for the step's render only
    assert abs(l_a - l_b) <= 1e-5 * abs(l_a)
    for a, c in zip(g_a, g_b):
        assert rel_l2(c, a) <= 1e-4
