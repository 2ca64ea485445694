"""CPU: host-side logic -- channels-last parameter storage with reference-compatible logical shapes / state-dict keys,
loading a reference-layout checkpoint and a Trainer checkpoint file, shard arithmetic, synthetic scene determinism, and the ray-sharded gradient
all-reduce on world_size = 2 (gloo)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from trinerflet_b200 import parallel, scene
from trinerflet_b200.network import NeRFNetwork
from trinerflet_b200.triplane_encoder import TriPlaneVolume, is_cl_coefs, is_cl_planes


def _net(C=16, R=128, S=4):
    return NeRFNetwork(bound=1.5, cuda_ray=True, density_thresh=10, triplane_channels=C, triplane_resolution=R,
                       triplane_wavelet_levels=S)


def test_state_dict_contract(golden_dir):
    net = _net()
    ref_keys = set(str(k) for k in np.load(os.path.join(golden_dir, "field_fp32.npz"))["state_dict_keys"])
    keys = set(net.state_dict().keys())
    assert ref_keys <= keys and keys - ref_keys == {"density_grid", "density_bitfield", "step_counter"}  # cuda_ray buffers
    enc = net.encoder
    assert enc.planes_features.shape == (3, 16, 32, 32) and is_cl_planes(enc.planes_features)
    assert [tuple(p.shape) for p in enc.planes_features_wavelet_coefs] == [(3, 16, 3, 32, 32), (3, 16, 3, 64, 64)]
    assert all(is_cl_coefs(p) for p in enc.planes_features_wavelet_coefs)
    assert all(float(p.detach().abs().sum()) == 0 for p in enc.planes_features_wavelet_coefs)   # zero-init (:220)
    assert net.cascade == 2 and net.density_grid.shape == (2, 128 ** 3) and net.density_bitfield.shape == (2 * 128 ** 3 // 8,)
    assert enc.output_dim == 48 and net.in_dim == 48 and net.sigma_net[0].weight.shape == (64, 48)
    assert net.color_net[0].weight.shape == (64, 31) and net.color_net[2].weight.shape == (3, 64)
    assert torch.equal(enc.plane_axes[0], torch.tensor([[1., 0.], [0., 0.], [0., 1.]]))     # 'up' plane = (x, z)


def test_load_reference_layout_checkpoint_keeps_channels_last():
    net = _net()
    sd = {k: v.clone().contiguous() for k, v in net.state_dict().items()}      # reference layout: NCHW-contiguous
    g = torch.Generator().manual_seed(0)
    sd["encoder.planes_features"] = torch.randn(3, 16, 32, 32, generator=g)
    sd["encoder.planes_features_wavelet_coefs.1"] = torch.randn(3, 16, 3, 64, 64, generator=g)
    sd.pop("encoder.planes_features_wavelet_coefs.0")                           # a finer level missing => stays zero (strict=False)
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert missing == ["encoder.planes_features_wavelet_coefs.0"] and not unexpected
    assert torch.equal(net.encoder.planes_features.detach(), sd["encoder.planes_features"])
    assert is_cl_planes(net.encoder.planes_features) and is_cl_coefs(net.encoder.planes_features_wavelet_coefs[1])
    opt = torch.optim.Adam(net.get_params(1e-2), betas=(0.9, 0.99), eps=1e-15)
    for p in net.parameters():
        p.grad = torch.ones_like(p)
    opt.step()
    assert is_cl_coefs(net.encoder.planes_features_wavelet_coefs[1])            # optimizer keeps the strides
    assert len(net.get_params(1e-2)) == 4


def test_trainer_checkpoint_file_round_trip(tmp_path):
    """the file Trainer.save_checkpoint(full=True) writes and the sequence Trainer.load_checkpoint reads it with
    (reconstruction/nerf/utils.py:1390-1430, :1466-1532): model state dict, mean_count / mean_density, optimizer and scaler
    state, through torch.save / torch.load; plus the 'best' flavour, which drops density_grid (:1450-1452)"""
    from trinerflet_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(3)
    src = _net(C=16, R=64, S=2)
    with torch.no_grad():
        for p in src.parameters():
            p.copy_(0.1 * torch.randn(p.shape, generator=g))
        src.density_grid.copy_(torch.rand(src.density_grid.shape, generator=g))
        src.density_bitfield.copy_(torch.randint(0, 256, src.density_bitfield.shape, generator=g, dtype=torch.uint8))
        src.step_counter.copy_(torch.randint(0, 4096, src.step_counter.shape, generator=g, dtype=torch.int32))
    src.mean_count, src.mean_density = 1234, 0.0625
    opt = torch.optim.Adam(src.get_params(1e-2), betas=(0.9, 0.99), eps=1e-15)
    for p in src.parameters():
        p.grad = torch.randn(p.shape, generator=g)
    opt.step()
    scaler_state = {"scale": 32768.0, "growth_factor": 2.0, "backoff_factor": 0.5, "growth_interval": 2000, "_growth_tracker": 7}
    state = {"epoch": 3, "global_step": 1200, "stats": {"loss": [0.1], "checkpoints": []},
             "mean_count": src.mean_count, "mean_density": src.mean_density,
             "optimizer": opt.state_dict(), "scaler": scaler_state, "model": src.state_dict()}
    path = tmp_path / "ngp_ep0003.pth"
    torch.save(state, path)

    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    dst = _net(C=16, R=64, S=2)
    gen0 = dst.bitfield_generation
    missing, unexpected = dst.load_state_dict(ckpt["model"], strict=False)          # :1482
    assert not missing and not unexpected
    dst.mean_count, dst.mean_density = ckpt["mean_count"], ckpt["mean_density"]     # :1491-1495
    assert dst.bitfield_generation > gen0                # a loaded occupancy grid invalidates work-lists built from the old one
    assert dst.mean_count == 1234 and dst.mean_density == 0.0625
    for (k, a), (_, b) in zip(src.state_dict().items(), dst.state_dict().items()):
        assert torch.equal(a, b), k
    assert is_cl_planes(dst.encoder.planes_features) and all(is_cl_coefs(p) for p in dst.encoder.planes_features_wavelet_coefs)
    # the optimizer the fast path uses accepts the torch.optim.Adam state of the file (:1512)
    fused = FusedAdam(dst.get_params(1e-2), betas=(0.9, 0.99), eps=1e-15)
    fused.load_state_dict(ckpt["optimizer"])
    st_src, st_dst = opt.state_dict()["state"], fused.state_dict()["state"]
    assert st_src.keys() == st_dst.keys()
    for k in st_src:
        assert torch.equal(st_src[k]["exp_avg"], st_dst[k]["exp_avg"]) and float(st_src[k]["step"]) == float(st_dst[k]["step"])

    best = dict(state, model={k: v for k, v in src.state_dict().items() if k != "density_grid"})
    torch.save(best, tmp_path / "best.pth")
    ckpt = torch.load(tmp_path / "best.pth", map_location="cpu", weights_only=False)
    dst = _net(C=16, R=64, S=2)
    missing, unexpected = dst.load_state_dict(ckpt["model"], strict=False)
    assert missing == ["density_grid"] and not unexpected
    assert torch.equal(dst.density_bitfield, src.density_bitfield) and float(dst.density_grid.abs().sum()) == 0.0


def test_unsupported_options_fail_loudly():
    with pytest.raises(NotImplementedError):
        TriPlaneVolume(number_of_features=16, plane_resolution=128, inner_multi_res_scale=4, learn_rotation_axis=True)
    with pytest.raises(NotImplementedError):
        NeRFNetwork(bound=1.5, cuda_ray=True, bg_radius=2.0, triplane_channels=16, triplane_resolution=128, triplane_wavelet_levels=4)
    with pytest.raises(ValueError):
        TriPlaneVolume(number_of_features=16, plane_resolution=128, inner_multi_res_scale=6)


def test_shards_and_scene():
    for n, w in [(60000, 8), (10, 3), (640000, 7), (5, 8)]:
        r = [parallel.shard_range(n, i, w) for i in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
    s1, s2 = scene.make_scene(), scene.make_scene()
    assert torch.equal(s1.poses, s2.poses) and abs(s1.intrinsics[0] - 1111.11) < 0.01
    assert torch.allclose(s1.poses[:, :3, 3].norm(dim=-1), torch.full((100,), scene.RADIUS), atol=1e-4)
    ro, rd, tgt = scene.sample_batch(s1, 1000, torch.Generator().manual_seed(0))
    assert torch.allclose(rd.norm(dim=-1), torch.ones(1000), atol=1e-5)
    assert float(((ro + 4 * rd).norm(dim=-1) < 1.6).float().mean()) > 0.5     # cameras look at the scene centre
    grid = scene.ball_density_grid(1.5, 0.75)
    occ = (grid > 0).float().mean(dim=1)
    assert abs(occ[0] - 4 / 3 * np.pi * 0.75 ** 3 / 8) < 0.01 and abs(occ[1] - 4 / 3 * np.pi * 0.75 ** 3 / 27) < 0.01
    assert scene.packbits_cpu(grid, 0.5).shape == (2 * 128 ** 3 // 8,)


def _ddp_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = _net(C=16, R=64, S=2)
    for i, p in enumerate(net.parameters()):
        p.grad = torch.full_like(p, float(rank + 1)) * (i + 1)
    parallel.allreduce_gradients(net, world)
    ok = all(torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1))) for i, p in enumerate(net.parameters()))
    local = torch.arange(*parallel.shard_range(11, rank, world)).float().unsqueeze(-1)
    full = parallel.gather_frame(local, 11, rank, world)
    ok = ok and torch.equal(full.squeeze(-1), torch.arange(11).float())
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_gradient_allreduce_and_gather_world2():
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_ddp_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]


def test_fused_adam_arithmetic_matches_torch_adam():
    """The update rule of csrc/optim.cu (k_adam_prepare + adam1), restated with torch ops, against torch.optim.Adam with
    the reference's hyper-parameters (main_nerf.py:119) and GradScaler-style unscaling -- pins the formula without a GPU."""
    import math
    torch.manual_seed(0)
    p_ref = torch.nn.Parameter(0.1 * torch.randn(257))
    opt = torch.optim.Adam([p_ref], lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    p = p_ref.detach().clone()
    m, v, step = torch.zeros_like(p), torch.zeros_like(p), 0.0
    b1, b2, lr, eps, scale = 0.9, 0.99, 1e-2, 1e-15, 65536.0
    for it in range(5):
        g_scaled = torch.randn(257) * scale * (0.0 if it == 2 else 1.0)      # one all-zero gradient step on the way
        p_ref.grad = g_scaled / scale
        opt.step()
        step += 1.0                                                            # k_adam_prepare
        bc1 = 1.0 - b1 ** step
        bc2s = math.sqrt(1.0 - b2 ** step)
        g = g_scaled * torch.tensor(1.0 / scale)                               # adam1
        m = m + (g - m) * (1.0 - b1)
        v = v * b2 + (1.0 - b2) * g * g
        p = p - (lr / bc1) * (m / (v.sqrt() / bc2s + eps))
        assert torch.allclose(p, p_ref.detach(), rtol=1e-5, atol=1e-8)
    from trinerflet_b200.optim import _dense_storage
    t = torch.zeros(3, 4, 4, 2).permute(0, 3, 1, 2)
    assert _dense_storage(t) and not _dense_storage(t[:, :1]) and not _dense_storage(torch.zeros(4, 4)[:, ::2])


def test_trunc_exp_matches_the_oracle_including_the_truncated_range():
    from oracle import field as of
    from trinerflet_b200.activation import trunc_exp
    x = torch.tensor([-40.0, -15.0001, -15.0, -3.25, 0.0, 0.5, 14.999, 15.0, 15.5, 30.0])
    g = torch.tensor([1.0, -2.0, 0.5, 3.0, 1.0, -1.0, 2.0, 1.0, 1.0, 0.25])
    a = x.clone().requires_grad_(True)
    b = x.clone().requires_grad_(True)
    ya, yb = trunc_exp(a), of.trunc_exp(b)
    ya.backward(g)
    yb.backward(g)
    assert torch.equal(ya, yb) and torch.equal(a.grad, b.grad)
    h = x.half().clone().requires_grad_(True)          # fp16 logits (autocast): fp32 value, gradient back in fp16
    y = trunc_exp(h)
    assert y.dtype == torch.float32
    y.backward(g)
    assert h.grad.dtype == torch.float16
