"""Timing of the two "next" rows built after the round-1 GPU budget was spent (run at the start of the next round):

  feeder   SURVEY.md 8f-2: rays.RayFeeder.select_batch (one launch, resident poses + images) against the reference-shaped
           path (rows of host ray tables -> pinned buffer -> H2D copy), per step of 60 000 rays on the Blender-shaped scene.
  render   SURVEY.md 8f-3: full-frame inference (800x800, max_steps as given) with the host-driven loop
           (model.infer_chunk = 0: one read-back per iteration) and the device-driven loop (infer_chunk = 4 / 8 / 16).

  sampling kernel-level timing of the point-ordered gather / scatter at base-light size.

  python profiles/bench_next_rows.py feeder|render|wide|sampling [--config base_light] [--max-steps 1024]

CUDA events around the whole operation, 3 warm-ups, median of 5; prints one JSON line per measurement."""
import argparse
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from trinerflet_b200 import rays, scene


def timed(fn, warm=3, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def bench_feeder(n_rays=60000, n_images=100):
    sc = scene.make_scene()
    H, W = scene.H_IMG, scene.W_IMG
    g = torch.Generator().manual_seed(0)
    images = torch.rand(n_images, H * W, 3, generator=g)                       # 768 MB fp32, as the reference keeps them
    feeder = rays.RayFeeder(sc.poses.cuda(), sc.intrinsics, H, W, images.cuda(), seed=0).shuffle()
    out = tuple(torch.empty(n_rays, 3, device="cuda") for _ in range(3))
    step = [0]

    def device_side():
        feeder.select_batch(step[0] % feeder.steps_per_epoch(n_rays), n_rays, out=out)
        step[0] += 1

    t_dev = timed(device_side, reps=20)
    # reference-shaped path: a pinned staging batch of (rays_o, rays_d, images) rows, copied to the device every step
    host = torch.empty(n_rays, 9).pin_memory()
    dev = torch.empty(n_rays, 9, device="cuda")
    t_h2d = timed(lambda: dev.copy_(host, non_blocking=True), reps=20)
    print(json.dumps({"what": "feeder", "rays": n_rays, "ms_device_feeder": t_dev, "ms_pinned_h2d_copy_only": t_h2d,
                      "note": "the reference additionally gathers the rows on the host from pageable tables (not timed here)"}))


def bench_render(config="base_light", max_steps=1024, chunks=(0, 4, 8, 16)):
    from trinerflet_b200.network import NeRFNetwork
    c = scene.CONFIGS[config]
    net = NeRFNetwork(bound=1.5, cuda_ray=True, density_thresh=10, min_near=0.2, triplane_channels=c["C"],
                      triplane_resolution=c["R"], triplane_wavelet_levels=c["S"], hidden_dim=c["hidden"],
                      hidden_dim_color=c["hidden"]).cuda()
    scene.init_model_(net, seed=0)
    scene.install_ball_occupancy(net, 0.75)
    net.eval()
    ro, rd = scene.full_frame(scene.make_scene(), 0)
    ro, rd = ro.cuda(), rd.cuda()
    ref = None
    for chunk in chunks:
        net.infer_chunk = chunk

        def frame():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
                return net.render(ro.unsqueeze(0), rd.unsqueeze(0), staged=True, bg_color=1, perturb=False, max_steps=max_steps)

        img = frame()["image"]
        if ref is None:
            ref = img
        ms = timed(frame, warm=2, reps=5)
        loop = getattr(net, "last_infer_loop", None) if chunk else None
        print(json.dumps({"what": "render", "config": config, "max_steps": max_steps, "infer_chunk": chunk, "ms_per_frame": ms,
                          "rays_per_s": ro.shape[0] / ms * 1e3, "max_abs_diff_vs_host_loop": float((img - ref).abs().max()),
                          "iterations": loop.iterations_done if loop else None, "state_reads": loop.reads if loop else None}))


def bench_sampling(C=32, R=2048, n_rays=60000):
    """kernel-level timing at base-light size: cell sort + point-ordered gather / scatter (+ tile-wise zero fill) on the samples
    a real march produces."""
    import numpy as np
    from trinerflet_b200 import _lib, raymarching as rm
    from trinerflet_b200._lib import call, ptr, stream
    from trinerflet_b200.idwt_plan import IdwtPlan
    from trinerflet_b200.network import NeRFNetwork
    from trinerflet_b200.triplane_encoder import cell_sort, cl_empty_planes
    net = NeRFNetwork(bound=1.5, cuda_ray=True, density_thresh=10, min_near=0.2, triplane_channels=C, triplane_resolution=R,
                      triplane_wavelet_levels=R // 64).cuda()
    scene.install_ball_occupancy(net, 0.75)
    plan = IdwtPlan.from_model(net)
    z = plan.zero
    sc = scene.make_scene()
    ro, rd, _ = scene.sample_batch(sc, n_rays, torch.Generator().manual_seed(0))
    ro, rd = ro.cuda(), rd.cuda()
    nears, fars = rm.near_far_from_aabb(ro, rd, net.aabb_train, 0.2)
    cnt = torch.zeros(2, dtype=torch.int32, device="cuda")
    xyzs, dirs, deltas, rays = rm.march_rays_train(ro, rd, 1.5, net.density_bitfield, 2, 128, nears, fars, cnt, -1, True, 128, True, 0, 1024)
    M = xyzs.shape[0]
    nv = cnt[0:1].clone()
    planes = cl_empty_planes(C, R, device="cuda").normal_().permute(0, 2, 3, 1)
    feat = torch.empty(M, 3 * C, device="cuda", dtype=torch.float16)
    gfeat = torch.randn(M, 3 * C, device="cuda").half()
    gpl = cl_empty_planes(C, R, device="cuda").permute(0, 2, 3, 1)
    inv = float(np.float32(1) / np.float32(1.5))
    perm = cell_sort(xyzs, 1.5, nv, 64)
    res = {"what": "sampling", "M": M, "C": C, "R": R}
    res["ms_cell_sort"] = timed(lambda: cell_sort(xyzs, 1.5, nv, 64))
    res["ms_point_fwd"] = timed(lambda: call("tnl_sample_planes_forward", ptr(planes), ptr(xyzs), M, R, C, inv, 1, ptr(nv), ptr(perm), ptr(feat), 1, stream()))
    res["ms_zero_fill"] = timed(lambda: plan.zero_gradient_tiles(gpl))
    res["ms_point_bwd"] = timed(lambda: call("tnl_sample_planes_backward", ptr(gfeat), 1, ptr(xyzs), M, R, C, inv, 1, ptr(nv), ptr(perm), ptr(gpl), stream()))
    print(json.dumps(res))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["feeder", "render", "sampling"])
    ap.add_argument("--config", default="base_light")
    ap.add_argument("--max-steps", type=int, default=1024)
    a = ap.parse_args()
    if a.what == "feeder":
        bench_feeder()
    elif a.what == "sampling":
        bench_sampling()
    else:
        bench_render(a.config, a.max_steps)
