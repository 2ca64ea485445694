"""GPU: the device-driven inference loop (tnl_infer_plan / tnl_march_rays_dev / tnl_composite_rays_dev /
tnl_compact_alive_dev, raymarching.DeviceRayLoop) against the host-driven loop of renderer.py:342-368 -- identical per-ray
results, fewer reads of the loop state.  The same comparison runs on the host emulation of the kernels in
tests/test_kernels_emu.py::test_device_driven_inference_loop_equals_host_driven.

(The file sorts after the other GPU test files on purpose: it was written after this round's GPU budget was spent, its
first run on hardware is the driver's round-end run.)"""
import numpy as np
import pytest
import torch

from tests.util import random_bitfield, synthetic_rays

pytestmark = pytest.mark.gpu
BOUND, CAS, H = 1.5, 2, 128


def _field(x, d):
    """a deterministic stand-in for the sigma / colour heads: a function of the sample alone"""
    s = (torch.sin(37.0 * x[:, 0] + 11.0 * x[:, 1]).abs() * 12.0)
    c = torch.cos(x * 5.0 + d).abs()
    return s, c.contiguous()


@pytest.mark.parametrize("max_steps,chunk,perturb", [(1024, 8, False), (48, 5, False), (1024, 3, True)])
def test_device_loop_equals_host_loop(max_steps, chunk, perturb):
    from trinerflet_b200 import raymarching as rm
    o, d = synthetic_rays(3000, 0)
    _, bits = random_bitfield(0)
    ro, rd, bf = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), bits.cuda()
    aabb = torch.tensor([-BOUND] * 3 + [BOUND] * 3, device="cuda")
    nears, fars = rm.near_far_from_aabb(ro, rd, aabb, 0.2)
    N = ro.shape[0]
    # host-driven (reference schedule)
    ws, dp, im = torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda"), torch.zeros(N, 3, device="cuda")
    alive, rt = torch.arange(N, dtype=torch.int32, device="cuda"), nears.clone()
    n_alive, step, iters = N, 0, 0
    torch.manual_seed(3)
    while step < max_steps and n_alive > 0:
        n_step = max(min(N // n_alive, 8), 1)
        x, dd, dl = rm.march_rays(n_alive, n_step, alive, rt, ro, rd, BOUND, bf, CAS, H, nears, fars, 128, perturb and step == 0,
                                  0, max_steps)
        s, c = _field(x, dd)
        rm.composite_rays(n_alive, n_step, alive, rt, s, c, dl, ws, dp, im, 1e-4)
        alive, cnt = rm.compact_rays_alive(alive, n_alive)
        n_alive = int(cnt.item())
        alive = alive[:n_alive]
        step += n_step
        iters += 1
    # device-driven
    ws2, dp2, im2 = torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda"), torch.zeros(N, 3, device="cuda")
    torch.manual_seed(3)      # the same torch.rand(N) draw for the first iteration's jitter
    loop = rm.DeviceRayLoop(ro, rd, nears, fars, BOUND, bf, CAS, H, 0, max_steps, perturb)
    loop.xyzs.fill_(7.0); loop.dirs.fill_(7.0); loop.deltas.fill_(7.0)     # dirty buffers: the marcher clears what it owns
    while True:
        for _ in range(chunk):
            x, dd = loop.begin_iteration()
            s, c = _field(x, dd)
            loop.end_iteration(s, c, ws2, dp2, im2, 1e-4)
        if loop.poll():
            break
    assert loop.iterations_done == iters and loop.reads < iters
    assert torch.equal(ws2, ws) and torch.equal(dp2, dp) and torch.equal(im2, im) and torch.equal(loop.rays_t, rt)


def test_render_with_device_loop_matches_host_loop():
    from tests.test_gpu_train import _model
    from trinerflet_b200 import scene
    net = _model("tiny")
    net.eval()
    sc = scene.make_scene()
    ro, rd = scene.full_frame(sc, 3)
    pick = torch.arange(0, ro.shape[0], 97)[:5000]
    ro, rd = ro[pick].cuda(), rd[pick].cuda()
    outs = []
    for chunk in (0, 8):
        net.infer_chunk = chunk
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            outs.append(net.render(ro.unsqueeze(0), rd.unsqueeze(0), staged=True, bg_color=1, perturb=False, max_steps=512))
    assert net.last_infer_loop.reads < net.last_infer_loop.iterations_done
    a, b = outs
    assert float(a["weights_sum"].sum()) > 10.0                                     # the ball is visible
    for k in ("image", "weights_sum"):
        assert (a[k].float() - b[k].float()).abs().max().item() <= 1e-6, k          # same kernels, same rows: expect 0
    m = torch.isfinite(a["depth"]) & torch.isfinite(b["depth"])
    assert (a["depth"][m] - b["depth"][m]).abs().max().item() <= 1e-6
