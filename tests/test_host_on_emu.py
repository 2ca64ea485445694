"""CPU: the package's HOST code (autograd Functions, NeRFRenderer.run_cuda / update_extra_state, IdwtPlan, RayFeeder,
DeviceRayLoop, PlaneGradReducer) running end to end on CPU tensors over the host build of the product kernels
(tests/emu_backend.py points the ctypes binding at tests/emu/_build/libkemu.so for the duration of a test), in fp32,
against the oracle pipeline.  The product has no CPU path; this is the test-suite executing the same Python + the same
kernel sources without a GPU.  The fused MLP kernels are not emulated: outside autocast NeRFNetwork runs the reference's
fp32 op sequence with torch, which is what these tests exercise."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from tests import emu_backend
from tests.util import rel_l2

BOUND = 1.5


@pytest.fixture
def emu(monkeypatch):
    return emu_backend.install(monkeypatch)


def _model(seed=0, radius=0.75, hidden=None):
    from trinerflet_b200 import scene
    from trinerflet_b200.network import NeRFNetwork
    c = scene.CONFIGS["tiny"]
    hidden = hidden or c["hidden"]
    net = NeRFNetwork(bound=BOUND, cuda_ray=True, density_thresh=10, min_near=0.2, triplane_channels=c["C"],
                      triplane_resolution=c["R"], triplane_wavelet_levels=c["S"], hidden_dim=hidden, hidden_dim_color=hidden)
    scene.init_model_(net, seed=seed)
    scene.install_ball_occupancy(net, radius)
    return net


def test_training_step_matches_oracle_pipeline_fp32(emu):
    """render (march -> sample -> fp32 heads -> composite) + MSE + wavelet regulariser, forward and backward, through
    NeRFNetwork.render and TrainStep.forward_backward vs oracle/pipeline.py::train_step on identical rays and jitter"""
    from oracle import pipeline
    from trinerflet_b200 import scene, trainer
    net = _model()
    net.train()
    sc = scene.make_scene()
    N = 384
    ro, rd, tgt = scene.sample_batch(sc, N, torch.Generator().manual_seed(0))
    opt = trainer.default_opt(fp16=False)
    ts = trainer.TrainStep(net, opt, None)
    torch.manual_seed(5)
    loss = ts.forward_backward(ro, rd, tgt, update_grid=False)
    M = int(net.step_counter[0, 0])
    assert M > 2000
    # oracle on the same parameters / rays / noises (the wrapper draws torch.rand(N) from the global generator)
    torch.manual_seed(5)
    noises = torch.rand(N).numpy()
    pf = net.encoder.planes_features.detach().clone().contiguous().requires_grad_(True)
    coefs = [p.detach().clone().contiguous().requires_grad_(True) for p in net.encoder.planes_features_wavelet_coefs]
    W = [w.detach().clone().requires_grad_(True) for w in net._weights()]
    loss_o, M_o = pipeline.train_step(pf, coefs, W, ro, rd, tgt, net.density_bitfield.numpy(), noises, lam=opt.wavelet_regularization)
    assert M_o == M
    assert abs(float(loss) - loss_o) <= 1e-5 * abs(loss_o)
    assert rel_l2(net.encoder.planes_features.grad, pf.grad) <= 1e-4
    for p, c in zip(net.encoder.planes_features_wavelet_coefs, coefs):
        assert rel_l2(p.grad, c.grad) <= 1e-4
    for w, wo in zip(net._weights(), W):
        assert rel_l2(w.grad, wo.grad) <= 1e-4
    # steady state (mean_count > 0: fixed buffer, no count read-back) gives the same loss
    net.mean_count = M
    net.zero_grad(set_to_none=True)
    torch.manual_seed(5)
    loss2 = ts.forward_backward(ro, rd, tgt, update_grid=False)
    assert abs(float(loss2) - float(loss)) <= 1e-6 * abs(float(loss))


def test_inference_render_host_loop_and_device_loop(emu):
    from trinerflet_b200 import parallel, scene
    net = _model()
    net.eval()
    sc = scene.make_scene()
    ro, rd = scene.full_frame(sc, 3)
    pick = torch.arange(0, ro.shape[0], 613)[:900]
    ro, rd = ro[pick].contiguous(), rd[pick].contiguous()
    outs = []
    for chunk in (0, 6):
        net.infer_chunk = chunk
        with torch.no_grad():
            outs.append(net.render(ro.unsqueeze(0), rd.unsqueeze(0), staged=True, bg_color=1, perturb=False, max_steps=256))
    loop = net.last_infer_loop
    assert loop.reads < loop.iterations_done and loop.iterations_issued >= loop.iterations_done
    a, b = outs
    assert float(a["weights_sum"].sum()) > 5.0
    assert torch.equal(a["image"], b["image"]) and torch.equal(a["weights_sum"], b["weights_sum"])
    m = torch.isfinite(a["depth"])
    assert torch.equal(a["depth"][m], b["depth"][m])
    # ray tiles are independent: shards rendered separately reproduce the frame (rendering partition)
    with torch.no_grad():
        parts = [net.render(ro[lo:hi].unsqueeze(0), rd[lo:hi].unsqueeze(0), staged=True, bg_color=1, perturb=False,
                            max_steps=256)["image"].view(-1, 3) for lo, hi in (parallel.shard_range(900, r, 3) for r in range(3))]
    assert (torch.cat(parts) - a["image"].view(-1, 3)).abs().max().item() <= 1e-4


def test_update_extra_state_and_plan_refresh(emu):
    from trinerflet_b200 import raymarching as rm
    from trinerflet_b200.idwt_plan import IdwtPlan
    net = _model()
    net.density_grid.zero_()
    net.iter_density = 0
    net.mean_density = 0
    net.update_extra_state()
    assert net.iter_density == 1 and float(net.density_grid.min()) >= 0 and net.mean_density > 0
    thresh = min(net.mean_density, net.density_thresh)
    assert torch.equal(net.density_bitfield, rm.packbits(net.density_grid, thresh))
    assert np.array_equal(net.density_bitfield.numpy(), np.packbits(net.density_grid.numpy().reshape(-1) > thresh, bitorder='little'))
    net.iter_density = 16
    before = net.density_grid.clone()
    net.update_extra_state()
    assert net.iter_density == 17 and bool((net.density_grid >= before * 0.95 - 1e-6).all())
    from trinerflet_b200 import scene
    scene.install_ball_occupancy(net, 0.4)
    plan = IdwtPlan.from_model(net)
    assert 0.0 < plan.stats["tile_fraction"] < 0.6
    assert plan.stats["zero_fill_fraction"] >= plan.stats["tile_fraction"]


def test_worklist_build_planes_autograd_equals_dense(emu):
    """triplane_encoder.build_planes_with_abs with an IdwtPlan (the autograd Function the training step uses) vs dense"""
    from tests.util import cl_coefs, cl_planes
    from trinerflet_b200.idwt_plan import IdwtPlan
    from trinerflet_b200.triplane_encoder import build_planes_with_abs
    C, n0, levels = 16, 16, 2
    R, T = n0 * 2 ** levels, n0 * 2 ** levels // 32
    g = torch.Generator().manual_seed(5)
    flags = torch.zeros(3, T, T, dtype=torch.bool)
    flags[:, 0, 1] = True
    flags[1, 1, 0] = True
    plan = IdwtPlan(R, n0, levels, C, "cpu").update(flags)
    pf = torch.randn(3, C, n0, n0, generator=g)
    coefs = [0.1 * torch.randn(3, C, 3, n0 * 2 ** l, n0 * 2 ** l, generator=g) for l in range(levels)]
    mask = flags.repeat_interleave(32, 1).repeat_interleave(32, 2)[:, None]
    gout = torch.randn(3, C, R, R, generator=g) * mask
    w_abs = torch.rand(levels, generator=g)
    res = []
    for p in (None, plan):
        pf_g = cl_planes(pf).requires_grad_(True)
        coefs_g = [cl_coefs(c).requires_grad_(True) for c in coefs]
        out, abs_sums = build_planes_with_abs(pf_g, coefs_g, p)
        ((out * gout).sum() + (abs_sums * w_abs).sum()).backward()
        res.append((out.detach(), abs_sums.detach(), pf_g.grad, [c.grad for c in coefs_g]))
    (o_d, a_d, gp_d, gc_d), (o_s, a_s, gp_s, gc_s) = res
    assert torch.equal(torch.where(mask, o_s, 0.0), torch.where(mask, o_d, 0.0))
    assert rel_l2(a_s, a_d) <= 1e-5
    assert torch.equal(gp_s, gp_d)
    for a, b in zip(gc_s, gc_d):
        assert torch.equal(a, b)


def test_ray_feeder_and_get_rays_host_code(emu, golden_dir):
    from oracle import rays as R
    from trinerflet_b200 import rays
    gold = np.load(os.path.join(golden_dir, "rays_ref.npz"))
    H, W, bs = 37, 53, int(gold["C_bs"])
    poses = torch.from_numpy(gold["B_poses"])
    images = torch.from_numpy(gold["C_images"]).view(4, H, W, 4)
    feeder = rays.RayFeeder(poses, gold["B_intr"], H, W, images)
    assert feeder.steps_per_epoch(bs) == 8
    feeder.shuffle(perm=torch.from_numpy(gold["C_perm"]))
    for b in (0, 3, 7):
        data = feeder.select_batch(b, bs)
        assert np.array_equal(data["rays_o"][0].numpy(), gold[f"C_b{b}_rays_o"])
        assert np.array_equal(data["rays_d"][0].numpy(), gold[f"C_b{b}_rays_d"])
        assert np.array_equal(data["images"][0].numpy(), gold[f"C_b{b}_images"])
    f0 = rays.RayFeeder(poses, gold["B_intr"], H, W, images, seed=3).shuffle()
    assert torch.equal(torch.sort(f0.perm).values, torch.arange(4 * H * W))
    whole = f0.select_batch(2, 1001)
    parts = [rays.RayFeeder(poses, gold["B_intr"], H, W, images, seed=3, rank=r, world_size=2).shuffle().select_batch(2, 1001)
             for r in range(2)]
    assert parts[0]["rays_d"].shape[1] == 501 and parts[1]["rays_d"].shape[1] == 500
    for k in ("rays_o", "rays_d", "images"):
        assert torch.equal(torch.cat([parts[0][k], parts[1][k]], 1), whole[k])
    bufs = tuple(torch.full((1001, w), float("nan")) for w in (3, 3, 4))
    f0.select_batch(2, 1001, out=bufs)
    assert torch.equal(bufs[1], whole["rays_d"][0]) and torch.equal(bufs[2], whole["images"][0])
    with pytest.raises(RuntimeError):
        f0.select_batch(2, 1001, out=tuple(torch.zeros(10, w) for w in (3, 3, 4)))

    # get_rays: the reference's signature, result dict and index-selection branches
    def check(res, n):
        assert res["rays_o"].shape == (4, n, 3) and res["rays_d"].shape == (4, n, 3) and res["inds"].shape == (4, n)
        inds = res["inds"].numpy()
        assert inds.min() >= 0 and inds.max() < H * W
        o, d, _ = R.get_rays_np(gold["B_poses"], gold["B_intr"], H, W, inds)
        assert np.array_equal(res["rays_o"].numpy(), o) and np.array_equal(res["rays_d"].numpy(), d)

    res = rays.get_rays(poses, gold["B_intr"], H, W, -1)
    check(res, H * W)
    assert np.array_equal(res["rays_d"].numpy(), gold["B_full_d"])
    torch.manual_seed(11)
    res = rays.get_rays(poses, gold["B_intr"], H, W, 257)                # the draw make_rays_golden.py recorded (utils.py:115)
    check(res, 257)
    assert np.array_equal(res["inds"].numpy(), gold["B_inds"]) and np.array_equal(res["rays_d"].numpy(), gold["B_rays_d"])
    check(rays.get_rays(poses, gold["B_intr"], H, W, 10 ** 9), H * W)
    res = rays.get_rays(poses, gold["B_intr"], H, W, 160, patch_size=4)
    check(res, 160)
    p = res["inds"][0].view(10, 16)
    assert torch.equal(p - p[:, :1], (torch.arange(4).view(4, 1) * W + torch.arange(4)).view(1, 16).expand(10, 16))
    res = rays.get_rays(poses, gold["B_intr"], H, W, 64, error_map=torch.rand(4, 128 * 128, generator=torch.Generator().manual_seed(1)))
    check(res, 64)
    assert res["inds_coarse"].shape == (4, 64)


# ------------------------------------------------------------------------------------------------ world_size 2 (gloo)
def _exchange_worker(rank, world, port, out):
    import torch.distributed as dist
    from _pytest.monkeypatch import MonkeyPatch
    mpatch = MonkeyPatch()
    try:
        emu_backend.install(mpatch)
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from tests.util import cl_planes
        from trinerflet_b200 import parallel
        net = _model(radius=0.5)
        R, C = net.encoder.plane_resolution, net.encoder.number_of_features
        ok = True
        for transport, tol in ((torch.float32, 0.0), (torch.bfloat16, 8e-3)):
            red = parallel.PlaneGradReducer(net, world, check=True, transport=transport).refresh()
            nt = R // red.tile
            flags = torch.zeros(3 * nt * nt, dtype=torch.bool)
            flags[red.tile_ids.long()] = True
            mask = flags.view(3, nt, nt).repeat_interleave(red.tile, 1).repeat_interleave(red.tile, 2)[:, None]
            gens = [torch.Generator().manual_seed(100 + r) for r in range(world)]
            grads = [torch.randn(3, C, R, R, generator=g) * mask for g in gens]          # what each rank's scatter produced
            mine = cl_planes(grads[rank].clone())
            red.reduce_(mine)
            want = sum(grads) / world
            err = ((mine - want).norm() / want.norm()).item()
            ok = ok and (torch.equal(mine, want) if tol == 0.0 else err <= tol) and 0 < red.fraction < 0.7
            # a gradient outside the dirty tiles must trip the debug check
            bad = cl_planes(torch.ones(3, C, R, R))
            try:
                red.reduce_(bad)
                ok = ok and bool(mask.all())
            except RuntimeError:
                pass
        out[rank] = bool(ok)
        dist.destroy_process_group()
    finally:
        mpatch.undo()


def test_sparse_plane_gradient_exchange_world2_gloo():
    """PlaneGradReducer (mark tiles -> pack -> all-reduce -> unpack/average) between two ranks over gloo, with the pack /
    unpack / mark kernels running on the host build: fp32 transport reproduces the dense average exactly, bf16 transport
    stays inside its stated rounding."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_exchange_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]


def _train_worker(rank, world, port, out, worklist=False):
    import torch.distributed as dist
    from _pytest.monkeypatch import MonkeyPatch
    mpatch = MonkeyPatch()
    try:
        emu_backend.install(mpatch)
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from trinerflet_b200 import parallel, scene, trainer
        N = 256
        sc = scene.make_scene()
        ro, rd, tgt = scene.sample_batch(sc, N, torch.Generator().manual_seed(0))
        opt = trainer.default_opt(fp16=False)
        # reference: the whole batch on one rank (no exchange)
        ref = _model()
        ref.train()
        torch.manual_seed(5)
        loss_ref = trainer.TrainStep(ref, opt, None).forward_backward(ro, rd, tgt, update_grid=False)
        # this rank's shard through the ray-sharded step (dirty-tile exchange of the plane gradient, fp32 transport)
        net = _model()
        net.train()
        lo, hi = parallel.shard_range(N, rank, world)
        ts = trainer.TrainStep(net, opt, None, world_size=world, check_sparse=not worklist, transport=torch.float32)
        ts.plan_on_any_device = worklist                 # with it: work-list IDWT step (partial zero fill of the gradient buffer)
        torch.manual_seed(5)
        torch.rand(lo)                                   # skip the jitter values of the rays before this shard
        loss = ts.forward_backward(ro[lo:hi], rd[lo:hi], tgt[lo:hi], update_grid=False)
        # shards have equal size: the average of the per-rank mean losses / gradients is the full-batch mean
        lt = loss.clone()
        dist.all_reduce(lt)
        ok = abs(float(lt) / world - float(loss_ref)) <= 1e-5 * abs(float(loss_ref))
        for (n, p), q in zip(net.named_parameters(), ref.parameters()):
            ok = ok and p.grad is not None and rel_l2(p.grad, q.grad) <= 2e-5
        out[rank] = bool(ok)
        dist.destroy_process_group()
    finally:
        mpatch.undo()


@pytest.mark.parametrize("worklist", [False, True])
def test_ray_sharded_training_step_world2_equals_single_rank(worklist):
    """two ranks, half of the rays each, replicated parameters; render backward -> dirty-tile exchange of the
    plane gradient (gloo here, NCCL on the box) -> IDWT backward; MLP gradients in one bucket.  Loss and every parameter
    gradient equal the single-rank step on the whole batch."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_train_worker, args=(2, port, out, worklist), nprocs=2, join=True)
    assert out[0] and out[1]


def _grid_refresh_worker(rank, world, port, out):
    import torch.distributed as dist
    from _pytest.monkeypatch import MonkeyPatch
    mpatch = MonkeyPatch()
    try:
        emu_backend.install(mpatch)
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from trinerflet_b200 import parallel, scene, trainer
        N = 301                                          # ragged shards: the ranks also draw different numbers of jitter values
        sc = scene.make_scene()
        ro, rd, tgt = scene.sample_batch(sc, N, torch.Generator().manual_seed(0))
        lo, hi = parallel.shard_range(N, rank, world)
        net = _model()
        net.train()
        ts = trainer.TrainStep(net, trainer.default_opt(fp16=False), None, world_size=world, transport=torch.float32)
        ts.plan_on_any_device = True
        torch.manual_seed(100 + rank)                    # every rank its own RNG stream, as bench.py seeds them
        l0 = ts.forward_backward(ro[lo:hi], rd[lo:hi], tgt[lo:hi], update_grid=False)      # steady-state step (builds the lists)
        gen0, tiles0 = net.bitfield_generation, ts.reducer.tile_ids.clone()
        net.zero_grad(set_to_none=True)
        l1 = ts.forward_backward(ro[lo:hi], rd[lo:hi], tgt[lo:hi], update_grid=True)       # dense step + update_extra_state
        net.zero_grad(set_to_none=True)
        l2 = ts.forward_backward(ro[lo:hi], rd[lo:hi], tgt[lo:hi], update_grid=False)      # work-list step on the NEW lists
        sig = torch.stack([net.density_bitfield.long().sum(), (net.density_bitfield.long() * torch.arange(net.density_bitfield.numel()) % 9973).sum(),
                           torch.tensor(ts.reducer.n_tiles), ts.reducer.tile_ids.long().sum()])
        both = [torch.zeros_like(sig) for _ in range(world)]
        dist.all_gather(both, sig)
        ok = all(torch.equal(b, both[0]) for b in both)                                    # identical occupancy + tile lists
        ok = ok and net.bitfield_generation > gen0 and ts._plan_gen == net.bitfield_generation == ts._reducer_gen
        ok = ok and not torch.equal(tiles0, ts.reducer.tile_ids) if tiles0.shape == ts.reducer.tile_ids.shape else ok
        ok = ok and all(bool(torch.isfinite(l)) for l in (l0, l1, l2))
        ok = ok and all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in net.parameters())
        # replicated parameters received identical (averaged) gradients
        for p in net.parameters():
            gs = [torch.zeros_like(p.grad.contiguous()) for _ in range(world)]
            dist.all_gather(gs, p.grad.contiguous())
            ok = ok and rel_l2(gs[1], gs[0]) <= 1e-6
        out[rank] = bool(ok)
        dist.destroy_process_group()
    finally:
        mpatch.undo()


def test_grid_refresh_keeps_ranks_consistent_world2_gloo():
    """update_extra_state runs on every rank from its own RNG stream; without synchronisation the density
    bitfields, hence the dirty-tile lists and the all-reduce buffer sizes, diverge.  Two gloo ranks, different seeds, ragged
    shards, update_grid=True: occupancy and tile lists stay identical, the next work-list step runs on the refreshed lists."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_grid_refresh_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]


def test_worklist_training_step_equals_dense_step(emu):
    """the optimised steady-state step (work-list IDWT forward, SplitIdwtBackward: clean part + active part, |yh| sums
    completed by the backward, gap lists) against the dense step on the same rays: same loss, same gradients"""
    from trinerflet_b200 import scene, trainer
    sc = scene.make_scene()
    N = 300
    ro, rd, tgt = scene.sample_batch(sc, N, torch.Generator().manual_seed(2))
    res = []
    for sparse in (False, True):
        net = _model(radius=0.45)
        net.train()
        ts = trainer.TrainStep(net, trainer.default_opt(fp16=False), None)
        ts.sparse_idwt = sparse
        ts.plan_on_any_device = True
        torch.manual_seed(9)
        loss = ts.forward_backward(ro, rd, tgt, update_grid=False)
        res.append((float(loss), [p.grad.clone() for p in net.parameters()], ts))
    (l_d, g_d, _), (l_s, g_s, ts) = res
    assert ts._plan is not None and 0 < ts._plan.stats["tile_fraction"] < 1
    assert abs(l_s - l_d) <= 1e-6 * abs(l_d)
    for a, b in zip(g_s, g_d):
        assert rel_l2(a, b) <= 1e-6


def test_bench_cpu_legs_and_kernel_attribution(capsys):
    """bench.py's host-only pieces: the reference arm prints the contract's JSON line from a directly timed, bounded sample;
    roofline attribution is per kernel (the work-list IDWT entry points split by their `parts` argument)."""
    import json
    import types
    import bench
    from trinerflet_b200 import scene
    assert bench.kernel_key("tnl_idwt_level_backward_sparse", (64, 32, 0.1, 128, 128, 1)) == "tnl_idwt_level_backward_sparse[active]"
    assert bench.kernel_key("tnl_idwt_level_backward_sparse", (64, 32, 0.1, 128, 128, 2)) == "tnl_idwt_level_backward_sparse[clean]"
    assert bench.kernel_key("tnl_mlp_backward", (1000,)) == "tnl_mlp_backward"
    assert bench.mlp_params(32, 64) == 13440 and bench.mlp_params(48, 128) == 41216
    args = types.SimpleNamespace(steps=1, warmup=0, gpus=1, config="tiny", scaling="weak")
    bench.run_reference(args, scene.CONFIGS["tiny"], 512)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "rays/s" and line["value"] > 0 and line["higher_is_better"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and "timed directly" in line["cpu_baseline"]["sample"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and abs(line["value"] - 512 / (line["ms_per_step"] * 1e-3)) <= 1e-6 * line["value"]


def test_mark_untrained_grid_matches_brute_force(emu):
    """NeRFRenderer.mark_untrained_grid (renderer.py:383-446) against a direct evaluation of the visibility rule over all
    128^3 x cascade cells (Morton order from the oracle)"""
    from oracle import raymarch as orc
    from trinerflet_b200 import scene
    net = _model()
    sc = scene.make_scene()
    poses = sc.poses[:3]
    net.density_grid.fill_(0.5)
    net.mark_untrained_grid(poses, sc.intrinsics)
    fx, fy, cx, cy = sc.intrinsics
    H = net.grid_size
    a = np.arange(H, dtype=np.int32)
    xx, yy, zz = np.meshgrid(a, a, a, indexing="ij")
    coords = np.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], -1)
    idx = orc.morton3D(coords).astype(np.int64)
    world = torch.from_numpy(2 * coords.astype(np.float32) / (H - 1) - 1)
    unseen_total = 0
    for cas in range(net.cascade):
        bound = min(2 ** cas, net.bound)
        hgs = bound / H
        pts = world * (bound - hgs)
        seen = torch.zeros(H ** 3, dtype=torch.bool)
        for b in range(poses.shape[0]):
            cam = (pts - poses[b, :3, 3]) @ poses[b, :3, :3]
            seen |= (cam[:, 2] > 0) & (cam[:, 0].abs() < cx / fx * cam[:, 2] + hgs * 2) & (cam[:, 1].abs() < cy / fy * cam[:, 2] + hgs * 2)
        want = np.full(H ** 3, 0.5, np.float32)
        want[idx[~seen.numpy()]] = -1.0
        assert np.array_equal(net.density_grid[cas].numpy(), want)
        unseen_total += int((~seen).sum())
    assert 0 < unseen_total < net.cascade * H ** 3
    # cells marked -1 stay out of the occupancy bitfield and are skipped by the EMA update (renderer.py:526-527)
    net.iter_density = 16
    net.update_extra_state()
    assert bool((net.density_grid[net.density_grid < 0] == -1).all()) and int((net.density_grid < 0).sum()) == unseen_total


def test_stage_growth_checkpoint_adds_a_zero_level(emu):
    """main_nerf.py:172-205 stages: (R, S) = (64, 2) -> (128, 4) keeps the base plane resolution, loads the coarser levels from the
    previous stage's checkpoint (strict=False, nerf/utils.py:1482) and adds one finer, zero-initialised level (SURVEY.md
    App. A-3).  The planes of the grown model equal the oracle's reconstruction from (old coefficients + a zero level)."""
    from oracle import wavelet as ow
    from trinerflet_b200.network import NeRFNetwork

    def make(R, S):
        return NeRFNetwork(bound=BOUND, cuda_ray=True, density_thresh=10, min_near=0.2, triplane_channels=8, triplane_resolution=R,
                           triplane_wavelet_levels=S)

    g = torch.Generator().manual_seed(0)
    s1 = make(64, 2)
    assert s1.encoder.planes_features.shape == (3, 8, 32, 32) and len(s1.encoder.planes_features_wavelet_coefs) == 1
    with torch.no_grad():
        s1.encoder.planes_features.copy_(torch.randn(3, 8, 32, 32, generator=g))
        s1.encoder.planes_features_wavelet_coefs[0].copy_(0.1 * torch.randn(3, 8, 3, 32, 32, generator=g))
    ckpt = {k: v.clone().contiguous() for k, v in s1.state_dict().items()}
    s2 = make(128, 4)
    assert s2.encoder.planes_features.shape == (3, 8, 32, 32) and len(s2.encoder.planes_features_wavelet_coefs) == 2
    missing, unexpected = s2.load_state_dict(ckpt, strict=False)
    assert missing == ["encoder.planes_features_wavelet_coefs.1"] and not unexpected
    assert float(s2.encoder.planes_features_wavelet_coefs[1].abs().sum()) == 0.0
    s2.encoder.reset_cahce()
    planes = s2.encoder.get_planes().detach()
    assert planes.shape == (3, 8, 128, 128)
    want = ow.build_planes(ckpt["encoder.planes_features"], [ckpt["encoder.planes_features_wavelet_coefs.0"], torch.zeros(3, 8, 3, 64, 64)])
    assert (planes - want).abs().max().item() <= 1e-5 * want.abs().max().item()


@pytest.mark.parametrize("C,half", [(32, False), (48, True)])
def test_field_mlp_function_wide_heads_forward(emu, C, half):
    """network._FieldMLP with the 128-wide heads of the "large" config on the host build: the forward (here the mma.sync
    kernels; the tcgen05 kernels of csrc/mlp_tc128.cu are GPU-only, tests/test_gpu_wide_heads.py) against the oracle's
    fp16-autocast emulation; the fused backward of the wide heads exists on the tensor cores only and says so"""
    from oracle import field as of
    from trinerflet_b200.network import _FieldMLP
    g = torch.Generator().manual_seed(C)
    M = 160
    W = of.init_mlp_weights(C, 128, 128, gen=g)
    feat = 0.5 * torch.randn(M, 3 * C, generator=g)
    d = torch.randn(M, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    s_o, rgb_o, _ = of.mlp_forward(feat, d, W, fp16=True)
    W_g = [w.clone().requires_grad_(True) for w in W]
    f_g = (feat.half() if half else feat.clone()).requires_grad_(True)
    nv = torch.tensor([M], dtype=torch.int32)
    s_g, rgb_g = _FieldMLP.apply(f_g, d, nv, *W_g)
    assert (rgb_g - rgb_o).abs().max().item() <= 2e-3 and rel_l2(s_g, s_o) <= 2e-3
    with pytest.raises(RuntimeError, match="128-wide heads"):
        (s_g.sum() + rgb_g.sum()).backward()


def test_fp16_training_configuration_matches_fp16_oracle(emu, monkeypatch, hidden=64):
    """the configuration every reference command trains with (--fp16): fp16 feature stream out of the sampler, fp16-rounded
    projected coordinates, the fused MLP kernels (here the mma.sync ones: the host build reports the tcgen05 kernels as
    unsupported) with their n_valid path -- against oracle/pipeline.py with its fp16-autocast emulation.  CUDA autocast
    cannot be switched on without a device, so the two queries the package makes about it are answered "fp16" for the test."""
    from oracle import pipeline
    from trinerflet_b200 import scene, trainer
    monkeypatch.setattr(torch, "is_autocast_enabled", lambda *a, **k: True)
    monkeypatch.setattr(torch, "get_autocast_dtype", lambda *a, **k: torch.float16)
    net = _model(hidden=hidden)
    net.train()
    sc = scene.make_scene()
    N = 320
    ro, rd, tgt = scene.sample_batch(sc, N, torch.Generator().manual_seed(3))
    opt = trainer.default_opt(fp16=False)            # (no real autocast context, no loss scaling: scale 1)
    ts = trainer.TrainStep(net, opt, None)
    torch.manual_seed(5)
    loss = ts.forward_backward(ro, rd, tgt, update_grid=False)
    M = int(net.step_counter[0, 0])
    torch.manual_seed(5)
    noises = torch.rand(N).numpy()
    pf = net.encoder.planes_features.detach().clone().contiguous().requires_grad_(True)
    coefs = [p.detach().clone().contiguous().requires_grad_(True) for p in net.encoder.planes_features_wavelet_coefs]
    W = [w.detach().clone().requires_grad_(True) for w in net._weights()]
    loss_o, M_o = pipeline.train_step(pf, coefs, W, ro, rd, tgt, net.density_bitfield.numpy(), noises, lam=opt.wavelet_regularization,
                                      fp16=True)
    assert M_o == M
    assert abs(float(loss) - loss_o) <= 2e-3 * abs(loss_o)
    assert rel_l2(net.encoder.planes_features.grad, pf.grad) <= 1e-2
    for p, c in zip(net.encoder.planes_features_wavelet_coefs, coefs):
        assert rel_l2(p.grad, c.grad) <= 1e-2
    for w, wo in zip(net._weights(), W):
        assert rel_l2(w.grad, wo.grad) <= 1e-2


def test_smoke_opt_in_section_runs(emu, monkeypatch, capsys):
    """__graft_entry__._smoke_opt_ins (the informational part of smoke()) on the host build: both lines report success"""
    import __graft_entry__ as ge
    from trinerflet_b200 import scene, trainer
    from trinerflet_b200.network import NeRFNetwork
    monkeypatch.setattr(torch.nn.Module, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    c = scene.CONFIGS["tiny"]
    sc = scene.make_scene()
    ro, rd, tgt = scene.sample_batch(sc, 256, torch.Generator().manual_seed(0))
    ge._smoke_opt_ins(scene, trainer, NeRFNetwork, c, sc, ro, rd, tgt)
    out = capsys.readouterr().out
    assert out.count("smoke opt-in") == 2 and "FAILED" not in out, out
    assert "rays_o equal: True" in out


def _render_worker(rank, world, port, out):
    import torch.distributed as dist
    from _pytest.monkeypatch import MonkeyPatch
    mpatch = MonkeyPatch()
    try:
        emu_backend.install(mpatch)
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from trinerflet_b200 import parallel, scene
        net = _model()
        net.eval()
        net.infer_chunk = 4 if rank == 1 else 0          # the two loops give identical pixels: mix them across ranks
        ro, rd = scene.full_frame(scene.make_scene(), 5)
        pick = torch.arange(0, ro.shape[0], 797)[:801]    # 801 rays: ragged shards (401 + 400)
        ro, rd = ro[pick].contiguous(), rd[pick].contiguous()
        with torch.no_grad():
            full = net.render(ro.unsqueeze(0), rd.unsqueeze(0), staged=True, bg_color=1, perturb=False, max_steps=256)
            net.infer_chunk = 4 if rank == 1 else 0
            frame = parallel.render_frame_sharded(net, ro, rd, rank, world, bg_color=1, max_steps=256)
        ok = frame["image"].shape == (801, 3) and frame["depth"].shape == (801,) and frame["weights_sum"].shape == (801,)
        ok = ok and (frame["image"] - full["image"].view(-1, 3)).abs().max().item() <= 1e-4
        ok = ok and (frame["weights_sum"] - full["weights_sum"].view(-1)).abs().max().item() <= 1e-4
        out[rank] = bool(ok)
        dist.destroy_process_group()
    finally:
        mpatch.undo()


def test_full_frame_render_sharded_world2_gloo():
    """BASELINE.json configs[4] / SURVEY.md 8e: a frame rendered as contiguous ray tiles on two ranks (no collective until the
    final gather of image / depth / weights_sum) equals the frame rendered by one rank; the shards are ragged"""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_render_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]


def test_fused_adam_resumes_from_torch_adam_checkpoint_host_logic(emu):
    """FusedAdam's host logic over the host build of csrc/optim.cu (the device run is tests/test_gpu_train.py): moments restored
    from a torch.optim.Adam checkpoint arrive NCHW-contiguous with a per-parameter `step`; the resumed update must pair every
    element with its own moments (channels-last parameters!) and continue the bias correction from the saved count, then keep
    tracking torch.optim.Adam step for step."""
    import copy
    from trinerflet_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(5)
    net_a, net_b = _model(), _model()
    net_b.load_state_dict(net_a.state_dict())
    adam = torch.optim.Adam(net_a.get_params(1e-2), betas=(0.9, 0.99), eps=1e-15)

    def grads():
        return [torch.randn(p.shape, generator=g) * (1 + i) for i, p in enumerate(net_a.parameters())]

    for _ in range(3):
        for p, gr in zip(net_a.parameters(), grads()):
            p.grad = gr
        adam.step()
    sd = copy.deepcopy(adam.state_dict())
    for st in sd["state"].values():
        for k in ("exp_avg", "exp_avg_sq"):
            st[k] = st[k].contiguous()                     # logical NCHW order, dense: NOT the channels-last storage order
    with torch.no_grad():
        for pa, pb in zip(net_a.parameters(), net_b.parameters()):
            pb.copy_(pa)
    fused = FusedAdam(net_b.get_params(1e-2), betas=(0.9, 0.99), eps=1e-15)
    fused.load_state_dict(sd)
    for k in range(2):
        for pa, pb, gr in zip(net_a.parameters(), net_b.parameters(), grads()):
            pa.grad, pb.grad = gr, gr.clone()
        adam.step()
        fused.step()
        assert float(fused.param_groups[0]["_tnl_state"][0]) == 4.0 + k
        for (n, pa), pb in zip(net_a.named_parameters(), net_b.parameters()):
            assert rel_l2(pb, pa) <= 1e-6, (n, k, rel_l2(pb, pa))
    # its own state dict round-trips too
    fused2 = FusedAdam(net_b.get_params(1e-2), betas=(0.9, 0.99), eps=1e-15)
    fused2.load_state_dict(copy.deepcopy(fused.state_dict()))
    for pa, pb, gr in zip(net_a.parameters(), net_b.parameters(), grads()):
        pa.grad, pb.grad = gr, gr.clone()
    adam.step()
    fused2.step()
    for (n, pa), pb in zip(net_a.named_parameters(), net_b.parameters()):
        assert rel_l2(pb, pa) <= 1e-6, n
